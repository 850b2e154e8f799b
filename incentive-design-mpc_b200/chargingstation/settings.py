"""Module-level knobs of the ``chargingstation`` package.  The names and values are the
reference's (``chargingstation/settings.py``: the mirrored classes import them by name); the
three ``*_SOLVER`` entries, cvxpy solver handles there, only name the CUDA kernels here."""

# verbosity of the closed loop and of the price solver: 0 silent .. 3 everything
PRINT_LEVEL = 1

# --- closed loop (charging_station.py) -------------------------------------------------
MIN_INITIAL_SOC = 0.3           # arriving EVs draw their state of charge from
MAX_INITIAL_SOC = 0.5           # U(MIN_INITIAL_SOC, MAX_INITIAL_SOC)
MIN_FULL_CHARGE_FRACTION = 0.95  # an EV leaves once its SoC exceeds this fraction of y_max
ADD_RESIDUAL_CHARGE_TO_BATTERY = False

# --- lower-level MPC (lompc.py): admissible constants -----------------------------------
MIN_MAX_BAT_SOC = 0.75          # y_max must lie in [MIN_MAX_BAT_SOC, MAX_MAX_BAT_SOC]
MAX_MAX_BAT_SOC = 0.9
MAX_BAT_CHARGE_RATE = 0.25      # w_max must lie in [0, MAX_BAT_CHARGE_RATE]
LOMPC_SOLVER = "B200_RICCATI_ACTIVE_SET"

# --- price loop (price_solver.py) ---------------------------------------------------------
MAX_PRICE_SOLVER_ITERATIONS = 1000
PRICE_SOLVER_TOL_TYPE = "avg"   # convergence test on the "avg" (mean) or the "max" EV error
PRICE_SOLVER_EPS_REG = 0.01     # proximal weight of the price step
PRICE_SOLVER_EPS_TOL = 0.01     # slack added to the tracking tolerance
PRICE_SOLVER_SOLVER = "B200_NNQP"

# --- upper-level MPC (bimpc.py) -------------------------------------------------------------
BIMPC_SOLVER = "B200_BLOCK_TRIDIAGONAL_IPM"
