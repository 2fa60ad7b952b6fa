"""UPDATE_PROFILE build only (TSD_NVCC_EXTRA=-DUPDATE_PROFILE): per-CTA cycle counters of k_update on the C2 workload."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ohm_tsd_slam_b200 import capi
from ohm_tsd_slam_b200.workload import DoubleLaserWorkload

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
wl = DoubleLaserWorkload(name, invert=capi.invert3x3, n_map=4 if name == "C3" else 6)
cfg = wl.cfg
g = capi.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
g.set_max_truncation(cfg.max_truncation)
wl.build_map(g)
g.set_timing(True)
L = capi.lib()
L.tsdg_debug_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
for filt in (0,):
    g.set_update_filter(filt)
    for st in wl.step_scans[:2]:
        g.stage_batch(list(st))
        for _ in range(3):
            g.push_staged()
        km = g.last_push_kernel_ms()
        buf = np.zeros((2048, 8), dtype=np.uint64)
        n = L.tsdg_debug_profile(g.h, buf.ctypes.data_as(C.c_void_p), 2048)
        b = buf[:n].astype(np.float64)
        us = 1.0 / 1965.0  # cycles -> us at the nominal clock
        print(f"filter {filt}: k_update {km['update'] * 1e3:.1f} us, {n} CTAs; per CTA (us): "
              f"producer total {b[:, 0].mean() * us:.1f} (max {b[:, 0].max() * us:.1f}) wait-empty {b[:, 1].mean() * us:.1f} grab {b[:, 2].mean() * us:.1f} "
              f"items {b[:, 3].mean():.1f} (min {b[:, 3].min():.0f} max {b[:, 3].max():.0f}) | consumer total {b[:, 4].mean() * us:.1f} (min {b[:, 4].min() * us:.1f} max {b[:, 4].max() * us:.1f}) "
              f"wait-full {b[:, 5].mean() * us:.1f} first-item {b[:, 6].mean() * us:.1f} | SMs used {len(np.unique(b[:, 7]))}")
g.set_update_filter(0)
