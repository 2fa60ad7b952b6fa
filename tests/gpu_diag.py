"""Diagnostic run on a GPU box: CUDA library vs the port oracle, verbose.  Not a pytest test.
usage: python tests/gpu_diag.py [config] [n_scans]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ohm_tsd_slam_b200 import capi, synth
from oracle import port
from tests.harness import Sequence, compare_grids

name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = synth.config(name)
print("devices", capi.device_count(), "invert equal", np.array_equal(capi.invert3x3(synth.pose_matrix(3, 2, 0.3)),
                                                                       port.invert3x3(synth.pose_matrix(3, 2, 0.3))))
seq = Sequence(cfg, port, capi, port.invert3x3)
scans = list(cfg.scans(n))
t = time.time()
ok, lines = seq.start(*scans[0])
print("start grid_equal", ok, lines, "stats", seq.ga.last_push_stats(), seq.gb.last_push_stats(), "%.1fs" % (time.time() - t))
for k, (pose, r) in enumerate(scans[1:]):
    t = time.time()
    out = seq.step(r)
    print(k, json.dumps({kk: vv for kk, vv in out.items()}, default=str), "%.1fs" % (time.time() - t))
# interpolation parity on random points
rng = np.random.default_rng(5)
xy = rng.uniform(-0.5, cfg.side + 0.5, size=(20000, 2))
ta, sa = seq.ga.interpolate_bilinear(xy)
tb, sb = seq.gb.interpolate_bilinear(xy)
print("interpolate status equal", np.array_equal(sa, sb), "tsd equal", np.array_equal(ta, tb, equal_nan=True), np.bincount(sa))
na, oa = seq.ga.interpolate_normal(xy)
nb, ob = seq.gb.interpolate_normal(xy)
print("normal ok equal", np.array_equal(oa, ob), "normals equal", np.array_equal(na[oa > 0], nb[ob > 0]))
print("launches", capi.kernel_launches())
