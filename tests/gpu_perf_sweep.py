"""Large-map push sweep (BASELINE.json configs[2]); not a pytest test: python tests/gpu_perf_sweep.py [layout_grid]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ohm_tsd_slam_b200.workload import large_grid_sweep

peak = None
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
out = large_grid_sweep(layout_grid=int(sys.argv[1]) if len(sys.argv) > 1 else 14, peak_gbs=peak)
print(out["grid"])
for r in out["rows"]:
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()})
