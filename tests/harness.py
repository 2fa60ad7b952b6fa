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
        self.icp_a.set_trace(True)
        self.icp_b.set_trace(True)
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


# ------------------------------------------------------------------------------------------------
# golden replay: the fixtures under tests/golden/ were produced by the reference itself
# (tests/make_golden.py); any backend with the port's surface can be replayed against them.
# ------------------------------------------------------------------------------------------------
import os

from ohm_tsd_slam_b200.scan import Scan

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def grid_checksum(g, states) -> np.ndarray:
    acc_t = np.uint64(0)
    acc_w = np.uint64(0)
    with np.errstate(over="ignore"):
        for p in np.nonzero(states == 2)[0]:
            t, w = g.download_partition(int(p))
            t = np.where(np.isnan(t), np.float64("nan"), t)
            acc_t += t.view(np.uint64).sum(dtype=np.uint64) * np.uint64(int(p) * 2 + 1)
            acc_w += w.view(np.uint64).sum(dtype=np.uint64) * np.uint64(int(p) * 2 + 1)
    return np.array([acc_t, acc_w], dtype=np.uint64)


def replay_golden_sequence(backend, name: str, exact_icp: bool, pose_tol: float = 1e-9, **kw):
    """Replays tests/golden/sequence_<name>.npz on `backend`; returns a list of failure strings."""
    G = np.load(os.path.join(GOLDEN, f"sequence_{name}.npz"))
    cfg = synth.config(name)
    fails = []
    g = backend.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid, **kw)
    g.set_max_truncation(cfg.max_truncation)
    icp = backend.Icp(30, 0.4, 0.02, g.bounds, **kw)
    icp.set_trace(True)
    n_scans = int(G["n_scans"])
    scans = list(cfg.scans(n_scans))
    # first scan: host-side sensor mirror (also checked against the reference's data/mask below)
    hs = HostSensor(cfg.sensor, lambda T: np.eye(3))
    (x, y, th), r0 = scans[0]
    hs.set_scan(r0)
    T0 = synth.pose_matrix(x, y, th)
    hs.transform(T0)
    inv = getattr(backend, "invert3x3")
    assert g.free_footprint(x, y, 0.6, 0.6)
    g.push(Scan(cfg.sensor, hs.data, hs.mask, hs.T, inv(hs.T)))
    for k in range(n_scans - 1):
        hs.set_scan(scans[k + 1][1])
        if not same(hs.data, G[f"data_{k}"]) or not same(hs.mask, G[f"mask_{k}"]):
            fails.append(f"scan {k}: host sensor data/mask differ from the reference's setStandardMask")
        pose = G[f"pose_{k}"]
        pinv = inv(pose)
        if not same(pinv, G[f"pose_inv_{k}"]):
            fails.append(f"scan {k}: invert3x3 differs from the reference's Matrix::invert")
        sc = Scan(cfg.sensor, G[f"data_{k}"], G[f"mask_{k}"], pose, G[f"pose_inv_{k}"])
        c, nrm, m, cnt = g.raycast_mask(sc, G[f"rays_{k}"])
        gm = G[f"rc_mask_{k}"]
        if not same(m, gm):
            fails.append(f"scan {k}: raycast mask differs at {int((m != gm).sum())} beams")
        both = (m > 0) & (gm > 0)
        if not same(c[both], G[f"rc_coords_{k}"][both]):
            fails.append(f"scan {k}: raycast coords differ, max abs {np.max(np.abs(c[both] - G[f'rc_coords_{k}'][both])):.3e}")
        if not same(nrm[both], G[f"rc_normals_{k}"][both]):
            fails.append(f"scan {k}: raycast normals differ")
        # ICP on the reference's own model/scene
        scene = np.zeros((cfg.sensor.beams, 2))
        valid = (~np.isinf(G[f"data_{k}"])) & (G[f"mask_{k}"] != 0)
        scene[valid, 0] = hs.rays_local[0, valid] * G[f"data_{k}"][valid]
        scene[valid, 1] = hs.rays_local[1, valid] * G[f"data_{k}"][valid]
        Mv = G[f"rc_coords_{k}"][gm > 0]
        Nv = G[f"rc_normals_{k}"][gm > 0]
        Sv = scene[valid]
        T, mse, pairs, its, st = icp.run(Mv, Nv, Sv, pose)
        gs = G[f"icp_stats_{k}"]
        if (pairs, its, st) != (int(gs[1]), int(gs[2]), int(gs[3])):
            fails.append(f"scan {k}: icp pairs/iterations/state {(pairs, its, st)} vs {tuple(gs[1:])}")
        if exact_icp:
            if not same(T, G[f"icp_T_{k}"]) or mse != gs[0]:
                fails.append(f"scan {k}: icp T/mse not bit-exact")
        else:
            if np.max(np.abs(T - G[f"icp_T_{k}"])) > pose_tol or abs(mse - gs[0]) > 1e-12:
                fails.append(f"scan {k}: icp T differs by {np.max(np.abs(T - G[f'icp_T_{k}'])):.3e}")
        cap = max(len(Mv), len(Sv), 1)
        nit, pm, ps, pc, _, _ = icp.trace(cap)
        gpc = G[f"icp_pair_count_{k}"]
        if nit != len(gpc) or not same(pc[:nit], gpc):
            fails.append(f"scan {k}: icp per-iteration pair counts differ")
        else:
            fm = np.concatenate([pm[i, :pc[i]] for i in range(nit)])
            fs = np.concatenate([ps[i, :pc[i]] for i in range(nit)])
            if not same(fm, G[f"icp_pairs_model_{k}"]) or not same(fs, G[f"icp_pairs_scene_{k}"]):
                fails.append(f"scan {k}: icp pair lists differ")
        # map update with the reference's pose
        pa = G[f"pose_after_{k}"]
        g.push(Scan(cfg.sensor, G[f"data_{k}"], G[f"mask_{k}"], pa, inv(pa)))
        st_, iw = g.partition_states()
        if not same(st_, G[f"states_{k}"]) or not same(iw, G[f"initw_{k}"]):
            fails.append(f"scan {k}: partition states / init weights differ")
    st_, _ = g.partition_states()
    for i, p in enumerate(G["dump_parts"]):
        t, w = g.download_partition(int(p))
        if not same(t, G["dump_tsd"][i]) or not same(w, G["dump_weight"][i]):
            fails.append(f"final cells of partition {int(p)} differ")
    if not same(grid_checksum(g, st_), G["checksum"]):
        fails.append("checksum over all cells differs")
    t, s_ = g.interpolate_bilinear(G["interp_xy"])
    if not same(s_, G["interp_status"]) or not same(t, G["interp_tsd"]):
        fails.append("interpolateBilinear differs")
    nn, ok = g.interpolate_normal(G["interp_xy"])
    if not same(ok, G["normal_ok"]) or not same(nn[ok > 0], G["normal_n"][ok > 0]):
        fails.append("interpolateNormal differs")
    return fails


def axis_map_scenario(cfg, n_scans: int, invert):
    """The scans of the map-publication fixtures (tests/golden/axis_map_*.npz): n_scans along the trajectory,
    pushed at their ground-truth poses."""
    hs = HostSensor(cfg.sensor, invert)
    out = []
    for pose, r in cfg.scans(n_scans):
        hs.set_scan(r)
        hs.T = synth.pose_matrix(*pose)
        out.append(hs.scan())
    return out


def check_axis_map(backend, name: str, **kw):
    """Replays tests/golden/axis_map_<name>.npz (generated from the reference) on `backend`: crossings, the quirky
    normals, the occupancy grid and the colour image must match bit for bit.  Returns failure strings."""
    G = np.load(os.path.join(GOLDEN, f"axis_map_{name}.npz"))
    cfg = synth.config(name)
    g = backend.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid, **kw)
    g.set_max_truncation(cfg.max_truncation)
    for sc in axis_map_scenario(cfg, int(G["n_scans"]), backend.invert3x3):
        g.push(sc)
    fails = []
    coords, normals, occ = g.axis_map(with_normals=True)
    if not same(coords, G["coords"]):
        fails.append(f"crossings differ: {coords.shape} vs {G['coords'].shape}")
    elif not same(normals, G["normals"]):
        fails.append("normals differ")
    if not same(occ, G["occupied"]):
        fails.append(f"occupancy differs in {int((occ != G['occupied']).sum())} cells")
    c2, n2, occ2 = g.axis_map(with_normals=False, occupied=np.full(occ.shape, 7, dtype=np.int8))
    if n2 is not None or not same(c2, G["coords"]):
        fails.append("crossings without normals differ")
    # cells the reference writes become 0 / -1, all others keep the caller's value
    go = G["occupied"]
    if not (same(occ2[go == 0], go[go == 0]) and np.isin(occ2[go == -1], (-1, 7)).all()):
        fails.append("occupancy (in/out array) differs")
    img = g.color_image(320, 200)
    if not same(img, G["image"]):
        fails.append(f"colour image differs in {int((img != G['image']).any(axis=2).sum())} pixels")
    return fails
