"""Exploration: what a push costs / updates on the C3 map for different obstacle sizes (not a pytest test)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ohm_tsd_slam_b200 import capi, synth
from ohm_tsd_slam_b200.workload import DoubleLaserWorkload
for scale in (0.5, 1.0, 2.0, 4.0):
    cfg0 = synth.config("C3")
    orig = synth.config
    def patched(which, _s=scale, _o=orig):
        c = _o(which)
        if which == "C3":
            c.obstacle_scale = _s
        return c
    synth.config = patched
    wl = DoubleLaserWorkload("C3", invert=capi.invert3x3, n_map=2, n_steps=2)
    synth.config = orig
    cfg = wl.cfg
    g = capi.Grid(cfg.cell_size, 5, cfg.layout_grid)
    g.set_max_truncation(cfg.max_truncation)
    wl.build_map(g)
    g.set_timing(True)
    for st in wl.step_scans[:1]:
        g.push_batch(list(st)); g.push_batch(list(st))
        s = g.last_push_stats(); km = g.last_push_kernel_ms()
        print(f"scale {scale}: active {s['active_tiles']} emptied {s['emptied_tiles']} updates {s['cell_updates']} k_update {km['update']*1e3:.1f} us classify {km['classify']*1e3:.1f} us -> {32*s['cell_updates']/km['update']/1e6/6543.1:.3f}")
    del g
