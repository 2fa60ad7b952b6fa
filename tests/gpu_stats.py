"""Push statistics of the bench workload (not a pytest test): python tests/gpu_stats.py [C2]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ohm_tsd_slam_b200 import capi
from ohm_tsd_slam_b200.workload import DoubleLaserWorkload

wl = DoubleLaserWorkload(sys.argv[1] if len(sys.argv) > 1 else "C2", invert=capi.invert3x3)
cfg = wl.cfg
g = capi.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
g.set_max_truncation(cfg.max_truncation)
for sc in wl.map_scans:
    g.push(sc)
g.fill(1.0, 1.0, only_uninitialized=True)
g.set_timing(True)
for step in wl.step_scans[:2]:
    for sc in step:
        g.push(sc)
        print(g.last_push_stats(), g.last_push_kernel_ms())
