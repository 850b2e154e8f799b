"""Generates tests/golden/log_schema.json: the `logs` dict schema of the reference's closed-loop simulator
(chargingstation/charging_station.py:118-149, parsed from its source - no import, cvxpy is not needed), the
`solver_stats` keys of PriceSolver.compute_optimal_prices (price_solver.py:167-173) and the keys the reference's
own consumers read (example/real_time_price_control_plots.py:24-305, plots/plots.py:115-127).

    python tests/golden/gen_log_schema.py [/root/reference]
"""
import ast
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def extract(ref_root: str) -> dict:
    cs = open(os.path.join(ref_root, "chargingstation", "charging_station.py")).read()
    tree = ast.parse(cs)
    produced = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "_init_logs":
            for st in node.body:
                if (isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Subscript)
                        and isinstance(st.targets[0].slice, ast.Constant)):
                    top = st.targets[0].slice.value
                    if isinstance(st.value, ast.Dict):
                        sub = {}
                        for k, v in zip(st.value.keys, st.value.values):
                            src = ast.get_source_segment(cs, v)
                            shape = re.search(r"zeros\(\s*\(([^)]*)\)", src)
                            dims = [d.strip().replace("self.", "") for d in shape.group(1).split(",") if d.strip()] if shape else []
                            sub[k.value] = {"dims": dims, "int": "dtype=int" in src or src.strip() == "0"}
                        produced[top] = sub
                    else:
                        produced[top] = None
    ps = open(os.path.join(ref_root, "chargingstation", "price_solver.py")).read()
    m = re.search(r"solver_stats\s*=\s*\{(.*?)\}", ps, re.S)
    stats_keys = re.findall(r"\"([a-z_]+)\"\s*:", m.group(1))
    consumed = set()
    for rel in ("chargingstation/example/real_time_price_control_plots.py", "chargingstation/plots/plots.py"):
        txt = open(os.path.join(ref_root, rel)).read()
        consumed |= {f"{a}.{b}" for a, b in re.findall(r"logs\[\"(\w+)\"\]\[\"(\w+)\"\]", txt)}
        consumed |= {a for a in re.findall(r"logs\[\"(\w+)\"\](?!\[)", txt)}
        consumed |= {f"solver_stats.{a}" for a in re.findall(r"stats\[\"(\w+)\"\]", txt)}
    return {"logs": produced, "solver_stats": stats_keys, "consumed": sorted(consumed)}


if __name__ == "__main__":
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    schema = extract(root)
    json.dump(schema, open(os.path.join(HERE, "log_schema.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(schema, indent=1)[:1500])
