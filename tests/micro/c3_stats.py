"""Exploratory: push statistics of the C3 bench workload (how many visited cells take the double-precision routes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ohm_tsd_slam_b200 import capi
from ohm_tsd_slam_b200.workload import DoubleLaserWorkload

wl = DoubleLaserWorkload("C3", invert=capi.invert3x3, n_map=4)
g = capi.Grid(wl.cfg.cell_size, wl.cfg.layout_partition, wl.cfg.layout_grid)
g.set_max_truncation(wl.cfg.max_truncation)
wl.build_map(g)
tot = {"cell_visits": 0, "cell_updates": 0, "fallback_cells": 0}
for st in wl.step_scans[:4]:
    for sc in st:
        g.push(sc)
        s = g.last_push_stats()
        for k in tot:
            tot[k] += s[k]
print(tot, "fallback share of visits %.4f" % (tot["fallback_cells"] / tot["cell_visits"]))
