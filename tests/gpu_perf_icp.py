"""Exploratory timing of the localisation path on a GPU box (not a pytest test): python tests/gpu_perf_icp.py [C1|C2|C3]
Ray cast from the current pose, then Icp::iterate (30 iterations) through the C ABI with host buffers; an
-DICP_PROFILE build of the library also prints k_icp's cycles per phase."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ohm_tsd_slam_b200 import capi
from ohm_tsd_slam_b200.workload import DoubleLaserWorkload

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
wl = DoubleLaserWorkload(name, invert=capi.invert3x3, n_map=4 if name == "C3" else 6)
cfg = wl.cfg
g = capi.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
g.set_max_truncation(cfg.max_truncation)
wl.build_map(g)
icp = capi.Icp(30, 0.4, 0.02, g.bounds)
sc0, rays0 = wl.step_scans[0][0], wl.step_rays[0][0]
hs = wl.sensors[0]
for _ in range(3):
    c, nrm, m, cnt = g.raycast_mask(sc0, rays0)
t0 = time.perf_counter()
for _ in range(20):
    c, nrm, m, cnt = g.raycast_mask(sc0, rays0)
rc = (time.perf_counter() - t0) / 20 * 1e3
valid = (~np.isinf(sc0.ranges)) & (sc0.mask != 0)
scene = np.ascontiguousarray(np.stack([hs.rays_local[0, valid] * sc0.ranges[valid], hs.rays_local[1, valid] * sc0.ranges[valid]], axis=1))
model, normals = np.ascontiguousarray(c[m > 0]), np.ascontiguousarray(nrm[m > 0])
for _ in range(3):
    out = icp.run(model, normals, scene, sc0.pose)
t0 = time.perf_counter()
for _ in range(20):
    out = icp.run(model, normals, scene, sc0.pose)
ic = (time.perf_counter() - t0) / 20 * 1e3
print(f"{name}: raycast {rc:.3f} ms ({cnt} hits), icp {ic:.3f} ms: model {len(model)} scene {len(scene)} -> "
      f"mse {out[1]:.3e} pairs {out[2]} iterations {out[3]} state {out[4]}")
