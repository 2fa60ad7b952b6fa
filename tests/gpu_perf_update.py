"""Exploratory timing of k_update on a GPU box (not a pytest test): python tests/gpu_perf_update.py [C2|C3] [reps]
The bench workload (double laser, dense regime), one batched launch pair per step: whole push, K2-only and
K3-only (tsdg_set_update_filter), as algorithmic GB/s (32 B per cell update) against the measured HBM peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ohm_tsd_slam_b200 import capi
from ohm_tsd_slam_b200.workload import DoubleLaserWorkload

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
peak = 6543.1
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
wl = DoubleLaserWorkload(name, invert=capi.invert3x3, n_map=4 if name == "C3" else 6)
cfg = wl.cfg
g = capi.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
g.set_max_truncation(cfg.max_truncation)
wl.build_map(g)
g.set_timing(True)
quick = len(sys.argv) > 3
for batch in ((2,) if quick else (2, 1)):
    for filt, label in ((0, "all"), (2, "K2 only"), (1, "K3 only")):
        rows = []
        for st in wl.step_scans:
            groups = [list(st)] if batch == 2 else [[s] for s in st]
            for gr in groups:
                g.set_update_filter(0)
                g.push_batch(gr)
                full = g.last_push_stats()
                g.set_update_filter(filt)
                g.push_batch(gr)
                part = g.last_push_stats()
                g.stage_batch(gr)
                ts = []
                for _ in range(reps):
                    g.push_staged()
                    ts.append(g.last_push_kernel_ms())
                upd = part["cell_updates"]
                u = float(np.median([t["update"] for t in ts]))
                c = float(np.median([t["classify"] for t in ts]))
                rows.append((upd, u, c, full["active_tiles"], full["emptied_tiles"]))
        g.set_update_filter(0)
        upd = np.mean([r[0] for r in rows]); u = np.mean([r[1] for r in rows]); c = np.mean([r[2] for r in rows])
        print(f"{name} batch={batch} {label:8s}: updates/launch {upd:.0f}  k_update {u * 1e3:.1f} us  k_classify {c * 1e3:.1f} us  "
              f"alg {32 * upd / u / 1e6:.0f} GB/s = {32 * upd / u / 1e6 / peak:.3f} of {peak:.0f}  "
              f"(active {np.mean([r[3] for r in rows]):.0f}, emptied {np.mean([r[4] for r in rows]):.0f} tiles)")
