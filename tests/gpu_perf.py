"""Exploratory timing on a GPU box (not a pytest test): python tests/gpu_perf.py [C2|C3] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ohm_tsd_slam_b200 import capi, synth
from ohm_tsd_slam_b200.scan import HostSensor

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = synth.config(name)
g = capi.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
g.set_max_truncation(cfg.max_truncation)
g.set_timing(True)
hs = HostSensor(cfg.sensor, capi.invert3x3)
(pose, r) = next(iter(cfg.scans(1)))
hs.set_scan(r)
hs.transform(synth.pose_matrix(*pose))
sc = hs.scan()
stream = torch.cuda.ExternalStream(g.stream_ptr)


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.sync()
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    g.sync()
    return e0.elapsed_time(e1) / n


# sparse regime: fresh map
t0 = time.time(); g.push(sc); print("first push (fresh map) wall %.2f ms" % ((time.time() - t0) * 1e3), g.last_push_stats(), g.last_push_kernel_ms())
for i in range(3):
    g.push(sc)
print("sparse regime push", g.last_push_stats(), g.last_push_kernel_ms())
g.stage_scan(sc)
ms = timed(g.push_staged, reps)
st = g.last_push_stats()
print("sparse staged push: %.3f ms/push  %.3f Gcell-updates/s" % (ms, st["cell_updates"] / ms / 1e6))

# dense regime
g.fill(0.5, 1.0)
for i in range(3):
    g.push(sc)
st = g.last_push_stats(); km = g.last_push_kernel_ms()
print("dense regime push", st, km)
g.stage_scan(sc)
ms = timed(g.push_staged, reps)
st = g.last_push_stats(); km = g.last_push_kernel_ms()
upd = st["cell_updates"]
print("dense staged push: %.3f ms/push  %.2f Gcell-updates/s  alg %.1f GB/s ; k_update %.3f ms -> %.1f GB/s (%.1f%% of 6543)" % (
    ms, upd / ms / 1e6, 32 * upd / ms / 1e6, km["update"], 32 * upd / km["update"] / 1e6, 100 * 32 * upd / km["update"] / 1e6 / 6543.1), km)
t0 = time.time()
for _ in range(reps):
    g.push(sc)
print("dense e2e push (host buffers, blocking): %.3f ms" % ((time.time() - t0) / reps * 1e3))

# raycast + icp
rays = hs.normalized_rays(cfg.cell_size).copy()
for _ in range(3):
    c, nrm, m, cnt = g.raycast_mask(sc, rays)
t0 = time.time()
for _ in range(reps):
    c, nrm, m, cnt = g.raycast_mask(sc, rays)
trc = (time.time() - t0) / reps
print("raycast e2e %.3f ms hits %d steps %s" % (trc * 1e3, cnt, g.raycast_steps()))
scene, ms_, _ = hs.scene()
icp = capi.Icp(30, 0.4, 0.02, g.bounds)
Mv, Nv, Sv = c[m > 0], nrm[m > 0], scene[ms_ > 0]
if len(Mv) > 2:
    for _ in range(3):
        out = icp.run(Mv, Nv, Sv, hs.pose)
    t0 = time.time()
    for _ in range(reps):
        out = icp.run(Mv, Nv, Sv, hs.pose)
    print("icp e2e %.3f ms" % ((time.time() - t0) / reps * 1e3), out[1:], len(Mv), len(Sv))
