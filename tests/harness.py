"""Shared parity harness: drives a grid / ICP / matcher backend through a synthetic scan sequence.

Backends with the same Python surface:
  oracle.port  -- plain-C restatement (CPU checker)
  ohm_tsd_slam_b200.capi -- the CUDA library through its C ABI (the product)
The reference itself (oracle.ref) has a different surface (it owns its SensorPolar2D) and is driven
separately in test_oracle_pin.py.
"""
from __future__ import annotations

import math

import numpy as np

from ohm_tsd_slam_b200 import synth
from ohm_tsd_slam_b200.scan import HostSensor


def same(a, b) -> bool:
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def compare_grids(ga, gb, max_report=3):
    """Bit-exact comparison of partition states, init weights and all cells (borders included).
    Returns (ok, report lines)."""
    sa, wa = ga.partition_states()
    sb, wb = gb.partition_states()
    lines = []
    ok = True
    if not same(sa, sb):
        bad = np.nonzero(sa != sb)[0]
        lines.append(f"partition state differs at {len(bad)} partitions, first {bad[:5]}: {sa[bad[:5]]} vs {sb[bad[:5]]}")
        ok = False
    if not same(wa, wb):
        bad = np.nonzero(wa != wb)[0]
        lines.append(f"initWeight differs at {len(bad)} partitions, first {bad[:5]}: {wa[bad[:5]]} vs {wb[bad[:5]]}")
        ok = False
    nbad = 0
    ncell = 0
    for p in np.nonzero((sa == 2) & (sb == 2))[0]:
        ta, wta = ga.download_partition(int(p))
        tb, wtb = gb.download_partition(int(p))
        dt = ~((ta == tb) | (np.isnan(ta) & np.isnan(tb)))
        dw = ~((wta == wtb) | (np.isnan(wta) & np.isnan(wtb)))
        if dt.any() or dw.any():
            nbad += 1
            ncell += int((dt | dw).sum())
            if nbad <= max_report:
                ys, xs = np.nonzero(dt | dw)
                y, x = int(ys[0]), int(xs[0])
                lines.append(f"partition {int(p)}: {int((dt | dw).sum())} cells differ, first (y={y},x={x}): "
                             f"tsd {ta[y, x]!r} vs {tb[y, x]!r}, w {wta[y, x]!r} vs {wtb[y, x]!r}; "
                             f"interior diffs {int((dt | dw)[:32, :32].sum())}")
    if nbad:
        lines.append(f"{nbad} partitions / {ncell} cells differ")
        ok = False
    return ok, lines


class Sequence:
    """A localisation + mapping loop over a synthetic scan sequence, run on backend A (the checker, which
    also decides the trajectory) and mirrored on backend B (the device under test)."""

    def __init__(self, cfg: synth.Config, backend_a, backend_b, invert, icp_iterations=30, dist=(0.4, 0.02),
                 b_kwargs=None):
        self.cfg = cfg
        self.A = backend_a
        self.B = backend_b
        b_kwargs = b_kwargs or {}
        self.ga = backend_a.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
        self.gb = backend_b.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid, **b_kwargs)
        self.ga.set_max_truncation(cfg.max_truncation)
        self.gb.set_max_truncation(cfg.max_truncation)
        self.sensor = HostSensor(cfg.sensor, invert)
        self.icp_a = backend_a.Icp(icp_iterations, dist[0], dist[1], self.ga.bounds)
        self.icp_b = backend_b.Icp(icp_iterations, dist[0], dist[1], self.gb.bounds, **b_kwargs)
        self.log = []

    def start(self, pose_xyt, ranges, footprint=(0.6, 0.6)):
        x, y, th = pose_xyt
        self.sensor.set_scan(ranges)
        self.sensor.transform(synth.pose_matrix(x, y, th))
        fa = self.ga.free_footprint(x, y, *footprint)
        fb = self.gb.free_footprint(x, y, *footprint)
        assert fa == fb
        sc = self.sensor.scan()
        self.ga.push(sc)
        self.gb.push(sc)
        return compare_grids(self.ga, self.gb)

    def step(self, ranges, check_grid=True):
        """One scan: raycast, ICP, pose update, push.  Returns a dict of comparison results."""
        cfg = self.cfg
        self.sensor.set_scan(ranges)
        sc = self.sensor.scan()
        rays = self.sensor.normalized_rays(cfg.cell_size).copy()
        ca, na, ma, cnta = self.ga.raycast_mask(sc, rays)
        cb, nb, mb, cntb = self.gb.raycast_mask(sc, rays)
        out = {}
        out["raycast_mask_equal"] = same(ma, mb)
        both = (ma > 0) & (mb > 0)
        out["raycast_coords_equal"] = same(ca[both], cb[both])
        out["raycast_normals_equal"] = same(na[both], nb[both])
        out["raycast_hits"] = (cnta, cntb)
        out["raycast_max_abs_diff"] = float(np.max(np.abs(ca[both] - cb[both]))) if both.any() else 0.0
        out["raycast_steps"] = (self.ga.raycast_steps(), self.gb.raycast_steps())
        scene, mask_s, _ = self.sensor.scene()
        Mv, Nv, Sv = ca[ma > 0], na[ma > 0], scene[mask_s > 0]
        Ta, msea, pa, ita, sta = self.icp_a.run(Mv, Nv, Sv, self.sensor.pose)
        Tb, mseb, pb, itb, stb = self.icp_b.run(Mv, Nv, Sv, self.sensor.pose)
        cap = max(len(Mv), len(Sv), 1)
        tra = self.icp_a.trace(cap)
        trb = self.icp_b.trace(cap)
        out["icp_T_max_abs_diff"] = float(np.max(np.abs(Ta - Tb)))
        out["icp_state"] = ((pa, ita, sta), (pb, itb, stb))
        out["icp_mse"] = (msea, mseb)
        n_it = min(tra[0], trb[0])
        pairs_equal = tra[0] == trb[0]
        first_bad = -1
        for it in range(n_it):
            ka, kb = tra[3][it], trb[3][it]
            if ka != kb or not same(tra[1][it, :ka], trb[1][it, :kb]) or not same(tra[2][it, :ka], trb[2][it, :kb]):
                pairs_equal = False
                first_bad = it
                break
        out["icp_pairs_equal"] = pairs_equal
        out["icp_first_bad_iteration"] = first_bad
        out["icp_iterations"] = (tra[0], trb[0])
        # the checker's pose drives both maps
        self.sensor.transform(Ta)
        sc2 = self.sensor.scan()
        self.ga.push(sc2)
        self.gb.push(sc2)
        out["push_stats"] = (self.ga.last_push_stats(), self.gb.last_push_stats())
        if check_grid:
            ok, lines = compare_grids(self.ga, self.gb)
            out["grid_equal"] = ok
            out["grid_report"] = lines
        self.last = dict(model=ca, mask_m=ma, scene=scene, mask_s=mask_s, normals=na)
        return out
