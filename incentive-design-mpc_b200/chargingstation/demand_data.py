"""Mirror of the reference's ``chargingstation/demand_data.py``: the 24-hour external
demand profile the example tiles over the simulation (demand_data.py:21-37).

The reference parses its bundled ``data/Real-Time Total Load.csv`` and uses only data rows
31-54, column 2: the 24 values of the "MediumTermLoadForecast" table (mid-hour load
forecast in MW, 18-Aug-2024).  Those 24 numbers are kept here as a constant so that the
example runs without the CSV; ``csv_path`` reads any file of the same layout instead.  (The
reference's ``main()`` only plots the two series with matplotlib, which this image lacks.)"""
from __future__ import annotations

import csv

import numpy as np

# MediumTermLoadForecast, Hour_End 1..24 (rows 31-54 of the reference's CSV)
MEDIUM_TERM_LOAD_FORECAST_MW = (
    73822, 70492, 69346, 67924, 67239, 67297, 67663, 69463, 72885, 77079, 80526, 84550,
    87982, 90588, 92603, 94458, 95772, 95887, 94438, 92268, 89947, 85908, 80634, 76068)


def _forecast_24(csv_path: str | None = None) -> np.ndarray:
    if csv_path is None:
        return np.asarray(MEDIUM_TERM_LOAD_FORECAST_MW, dtype=float)
    with open(csv_path, newline="") as csvfile:
        data = list(csv.reader(csvfile, delimiter=","))
    return np.asarray(data[30:54]).astype(float)[:, 1]  # demand_data.py:26


def medium_term_demand_forecast(hours: int, scale: float, interpolate: bool = False,
                                csv_path: str | None = None) -> np.ndarray:
    """demand_data.py:21-37: hourly (or, with ``interpolate``, half-hourly) demand for
    ``hours`` hours, the 24-hour forecast repeated, times ``scale``."""
    f24 = _forecast_24(csv_path)
    # Interpolated demand forecasts every 30 mins, starting from 00:00.
    f48 = np.zeros((48,))
    f48[1::2] = f24
    f48[0::2] = (f24 + np.roll(f24, 1)) / 2
    f48_ = f48.tolist()
    demand = f48_ * (hours // 24) + f48_[: 2 * (hours % 24)]
    if not interpolate:
        demand = demand[0::2]
    return scale * np.array(demand)
