"""Benchmark workload of bench.py: BASELINE.json configs[1] -- the double-laser configuration (two 1081-beam
scanners with local offsets +-0.35 m sharing one 4096 x 4096 grid, reference src/SlamNode.cpp:104-121), in the
DENSE regime (every partition already allocated, BASELINE.md section 2 "observation"): the map is first built
by pushing scans along the trajectory, then every partition that is still unallocated is allocated as free
space, so that each timed push rewrites every in-range visible cell (the bandwidth-bound regime the north
star's HBM target refers to).  Plain numpy; the same scans feed the CUDA arm and the reference arm."""
from __future__ import annotations

import math

import numpy as np

from . import synth
from .scan import HostSensor


class DoubleLaserWorkload:
    def __init__(self, config_name: str = "C2", n_map: int = 6, n_steps: int = 4, invert=None, seed_offset: int = 0,
                 origin=None):
        """origin: (x, y) of the robot's start and of its room's centre; default = the grid centre."""
        self.cfg = cfg = synth.config(config_name)
        self.name = config_name
        self.offsets = (+0.35, -0.35)
        rng = np.random.default_rng(cfg.seed + 7 + seed_offset)
        traj = cfg.trajectory(n_map + n_steps)
        if origin is None:
            self.world = cfg.world()
        else:
            self.world = synth.World.room_with_obstacles(origin[0], origin[1], cfg.room[0], cfg.room[1], cfg.n_obstacles,
                                                         cfg.seed + seed_offset, obstacle_scale=cfg.obstacle_scale)
            traj = traj + np.array([origin[0] - cfg.side / 2.0, origin[1] - cfg.side / 2.0, 0.0])
        self.sensors = [HostSensor(cfg.sensor, invert) for _ in self.offsets]
        self.map_scans = []    # list of Scan (alternating lasers)
        self.step_scans = []   # list of (ScanA, ScanB)
        self.step_rays = []
        for i, (x, y, th) in enumerate(traj):
            pair = []
            rays = []
            for hs, off in zip(self.sensors, self.offsets):
                sx = x - off * math.sin(th)
                sy = y + off * math.cos(th)
                r = synth.scan_from_pose(self.world, cfg.sensor, sx, sy, th, rng)
                hs.set_scan(r)
                hs.T = np.eye(3)
                hs.rays = hs.rays_local.copy()
                hs.ray_norm = 1.0
                hs.transform(synth.pose_matrix(sx, sy, th))
                pair.append(hs.scan())
                rays.append(hs.normalized_rays(cfg.cell_size).copy())
            if i < n_map:
                self.map_scans.extend(pair)
            else:
                self.step_scans.append(tuple(pair))
                self.step_rays.append(tuple(rays))

    def describe(self) -> dict:
        c = self.cfg
        return {
            "workload": f"{self.name}: double-laser (2 x {c.sensor.beams} beams, 270 deg, max range {c.sensor.max_range:g} m) on a "
                        f"{c.cells}x{c.cells} TsdGrid @ {c.cell_size * 100:g} cm, 32x32 partitions, dense regime "
                        f"(all {(c.cells // 32) ** 2} partitions allocated), room {c.room[0]:g}x{c.room[1]:g} m + {c.n_obstacles} obstacles",
            "grid_cells": c.cells * c.cells,
            "cell_state_bytes": (c.cells // 32) ** 2 * 8832 * 2,
            "pushes_per_step": 2,
            "l2_policy": f"cell state ({(c.cells // 32) ** 2 * 8832 * 2 / 1e6:.0f} MB) is larger than L2 (126 MB): inputs larger than L2, no flush",
        }

    def build_map(self, grid):
        """Same sequence on any backend with .push/.fill: build, then allocate the rest as free space."""
        for sc in self.map_scans:
            grid.push(sc)
        grid.fill(1.0, 1.0, only_uninitialized=True)


class MultiRobotWorkload:
    """bench.py at N > 1 GPUs (BASELINE.json configs[4], "multi-SLAM" on a grid sharded in bands): N double-laser
    robots of the C2 kind, each with its own room, on ONE TsdGrid of (8192 N)^2 cells sharded in N bands of 8192
    cell rows (256 partition rows; N = 8 is the 65536^2 grid of configs[4]).  Robot k starts 20 m above the lower
    edge of band k, so its room (67.5 m tall) straddles the boundary to band k-1: the lower 14 m of every room
    are integrated by the neighbouring GPU and the boundary rows really change every step.  Per-GPU work is one
    robot's worth for every N: weak scaling."""

    BAND_ROWS = 8192
    RISE = 20.0  # metres between a band's lower edge and its robot's start

    def __init__(self, n_robots: int, config_name: str = "C2", n_map: int = 6, n_steps: int = 4, invert=None):
        base = synth.config(config_name)
        self.n = n_robots
        cells = self.BAND_ROWS * n_robots
        self.layout_grid = int(round(math.log2(cells)))
        if (1 << self.layout_grid) != cells:
            raise ValueError("the number of GPUs must be a power of two (the reference's grids are 2^k cells wide)")
        self.cell_size = base.cell_size
        self.max_truncation = base.max_truncation
        side = cells * base.cell_size
        band_h = self.BAND_ROWS * base.cell_size
        self.robots = [DoubleLaserWorkload(config_name, n_map, n_steps, invert, seed_offset=1000 * k,
                                           origin=(side / 2.0, k * band_h + self.RISE)) for k in range(n_robots)]
        self.cfg = base
        self.sensors = self.robots[0].sensors
        # one step = every robot's two scans, robot after robot
        self.map_scans = [sc for i in range(0, 2 * n_map, 2) for r in self.robots for sc in r.map_scans[i:i + 2]]
        self.step_scans = [tuple(sc for r in self.robots for sc in r.step_scans[i]) for i in range(n_steps)]
        self.step_rays = [tuple(ry for r in self.robots for ry in r.step_rays[i]) for i in range(n_steps)]

    def describe(self) -> dict:
        c = self.cfg
        cells = 1 << self.layout_grid
        band_bytes = (self.BAND_ROWS // 32 + 2) * (cells // 32) * 8832 * 2
        return {
            "workload": f"{self.n} x C2 robot (double-laser, 2 x {c.sensor.beams} beams, max range {c.sensor.max_range:g} m, room "
                        f"{c.room[0]:g}x{c.room[1]:g} m + {c.n_obstacles} obstacles each) on one {cells}x{cells} TsdGrid @ "
                        f"{c.cell_size * 100:g} cm sharded in {self.n} bands of {self.BAND_ROWS} cell rows, dense regime; every room "
                        f"straddles a band boundary",
            "grid_cells": cells * cells,
            "cell_state_bytes_per_gpu": band_bytes,
            "pushes_per_step": 2 * self.n,
            "l2_policy": f"cell state per GPU ({band_bytes / 1e9:.1f} GB) is larger than L2 (126 MB): inputs larger than L2, no flush",
        }


class HypothesisWorkload:
    """BASELINE.json configs[3] (SURVEY.md 8d "C4"): one C1 scan, 10^5 pose hypotheses (model index, scene index),
    a control set of 360 scene points, for the three scorers.  The matcher's pre-processing stays on the host in the
    reference (RandomMatching.cpp:41-183); here the inputs are built with plain numpy: orientations by central
    differences over the beam neighbours (the scorers only need *an* orientation per point), control set = the
    first 360 valid scene points at a stride, hypotheses = all valid (model, scene) index pairs inside the angular
    window, in canonical order, truncated / repeated to exactly n_hyp."""

    PDF_PARAMS = np.array([0.45, 0.0, 0.25, 0.05, 0.25, 0.9, 20.0, np.pi / 180.0 * 3, 0.2, 0.08, 3.0, 0.5])  # PDFMatching ctor defaults

    def __init__(self, model_xy, model_mask, scene_xy, scene_mask, spec, n_hyp: int = 100000, n_control: int = 360,
                 phi_max: float = math.radians(30.0)):
        M, S = np.asarray(model_xy, dtype=np.float64), np.asarray(scene_xy, dtype=np.float64)
        n = len(M)
        assert len(S) == n
        self.M, self.S = M, S

        def orientation(P, mask):
            d = np.zeros_like(P)
            d[1:-1] = P[2:] - P[:-2]
            ok = np.zeros(n, dtype=bool)
            ok[1:-1] = (mask[2:] > 0) & (mask[:-2] > 0) & (mask[1:-1] > 0)
            # normal = tangent rotated by -90 deg; its angle
            return np.where(ok, np.arctan2(-d[:, 0], d[:, 1]), 0.0), ok

        self.phi_m, ok_m = orientation(M, np.asarray(model_mask))
        self.phi_s, ok_s = orientation(S, np.asarray(scene_mask))
        self.idx_m_valid = np.nonzero(ok_m)[0].astype(np.int32)
        idx_s = np.nonzero(ok_s)[0]
        stride = max(1, len(idx_s) // n_control)
        self.idx_control = idx_s[::stride][:n_control].astype(np.int32)
        C = len(self.idx_control)
        self.control = np.ones((3, C))
        self.control[0] = S[self.idx_control, 0]
        self.control[1] = S[self.idx_control, 1]
        self.phi_control = self.phi_s[self.idx_control]
        self.phi_max = phi_max
        span = int(phi_max / spec.angular_res)
        pairs = []
        for im in self.idx_m_valid:
            lo, hi = max(0, im - span), min(n - 1, im + span)
            js = idx_s[(idx_s >= lo) & (idx_s <= hi)]
            pairs.append(np.stack([np.full(len(js), im), js], axis=1))
            if sum(len(p) for p in pairs) >= n_hyp:
                break
        h = np.concatenate(pairs).astype(np.int32)
        reps = -(-n_hyp // len(h))
        self.hyps = np.ascontiguousarray(np.tile(h, (reps, 1))[:n_hyp])
        self.model_valid = M[self.idx_m_valid]
        self.phi_valid = self.phi_m[self.idx_m_valid]
        self.model_angles = np.arctan2(self.model_valid[:, 1], self.model_valid[:, 0])
        self.model_dists = np.hypot(self.model_valid[:, 0], self.model_valid[:, 1])
        self.theta_min, self.theta_max = spec.phi_min, spec.phi_min + spec.angular_res * (spec.beams - 1)

    def run_tsd(self, matcher, grid, pose, hyps=None):
        return matcher.score_tsd(grid, self.hyps if hyps is None else hyps, self.M, self.S, self.phi_m, self.phi_s, self.phi_max,
                                 self.control, pose, 0.25)

    def run_rnm(self, matcher, hyps=None):
        return matcher.score_rnm(self.hyps if hyps is None else hyps, self.M, self.S, self.phi_m, self.phi_s, self.phi_max,
                                 self.control, self.phi_control, self.model_valid, self.phi_valid, self.theta_min,
                                 self.theta_max, 1.0 / 0.15 ** 2, 0.33, self.control.shape[1] // 3)

    def run_pdf(self, matcher, hyps=None):
        return matcher.score_pdf(self.hyps if hyps is None else hyps, self.M, self.S, self.phi_m, self.phi_s, self.phi_max,
                                 self.control, self.model_angles, self.model_dists, self.PDF_PARAMS)


def hypothesis_benchmark(device: int = 0, n_hyp: int = 100000, reps: int = 3, dist=None, keep_inputs: bool = False):
    """C4 on the device: a C1 map, one scan, n_hyp hypotheses through the three scorers (C ABI, host buffers:
    H2D of hypotheses + control set and D2H of the per-hypothesis results inside the timing).  With `dist`
    (torch.distributed, one rank per GPU) the hypothesis list is split in contiguous slices and the winner merged
    (sharded.merge_best_hypothesis); times are the max over ranks.  keep_inputs: also return the inputs under
    "_inputs" (bench.py's CPU baseline leg times a checker on a sample of them)."""
    import time

    from . import capi
    from .sharded import merge_best_hypothesis

    cfg = synth.config("C1")
    g = capi.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid, device=device)
    g.set_max_truncation(cfg.max_truncation)
    hs = HostSensor(cfg.sensor, capi.invert3x3)
    scans = list(cfg.scans(4))
    for pose, r in scans[:3]:
        hs.set_scan(r)
        hs.T = np.eye(3)
        hs.rays = hs.rays_local.copy()
        hs.ray_norm = 1.0
        hs.transform(synth.pose_matrix(*pose))
        g.push(hs.scan())
    pose, r = scans[3]
    hs.set_scan(r)
    hs.T = np.eye(3)
    hs.rays = hs.rays_local.copy()
    hs.ray_norm = 1.0
    hs.transform(synth.pose_matrix(*pose))
    sc = hs.scan()
    M, _, mM, _ = g.raycast_mask(sc, hs.normalized_rays(cfg.cell_size).copy())
    S, mS, _ = hs.scene()
    wl = HypothesisWorkload(M, mM, S, mS, cfg.sensor, n_hyp=n_hyp)
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    lo, hi = rank * n_hyp // world, (rank + 1) * n_hyp // world
    mine = np.ascontiguousarray(wl.hyps[lo:hi])
    mt = capi.Matcher(device)
    out = {"n_hypotheses": n_hyp, "control_points": int(wl.control.shape[1]), "model_points_valid": int(len(wl.model_valid)),
           "n_gpus": world}

    dev_words = {}

    def amax(t):
        # one 8-byte word through a device tensor that lives for the whole benchmark (NCCL reduces device memory)
        import torch
        w = dev_words.get(t.dtype)
        if w is None:
            w = dev_words[t.dtype] = torch.zeros(1, dtype=t.dtype, device=torch.device("cuda", device))
        w.copy_(t)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        t.copy_(w)

    def merged(name, res):
        score = res[0] if name != "rnm" else res[2]
        b = res[1] if name == "tsd" else (res[3] if name == "rnm" else res[2])
        return merge_best_hypothesis(float(max(score[b], 0.0)) if b >= 0 else 0.0, lo + b if b >= 0 else -1, amax)

    runs = {"tsd": lambda h: wl.run_tsd(mt, g, hs.pose, h), "rnm": lambda h: wl.run_rnm(mt, h), "pdf": lambda h: wl.run_pdf(mt, h)}
    for name, fn in runs.items():
        for _ in range(2):  # warm-up: the scorer AND the merge (the first collective pays NCCL's lazy set-up)
            res = fn(mine)
            if dist:
                merged(name, res)
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            res = fn(mine)
            if dist:
                merged(name, res)
        dt = (time.perf_counter() - t0) / reps
        if dist:
            import torch
            t = torch.tensor([dt], dtype=torch.float64)
            amax(t)
            dt = float(t[0])
        out[name] = {"ms": dt * 1e3, "hypotheses_per_s": n_hyp / dt}
    if world == 1:
        # the whole relocalisation from the raw model / scene points -- pre-processing on the device
        # (match_prepare: normals, subsampling, control set, trials, hypothesis list) + scoring of the resident set
        phi_max = math.radians(30.0)
        leg = {}
        for name in ("tsd", "rnm", "pdf"):
            kw = ({"grid": g, "t_sensor": hs.pose, "zrand": 0.25} if name == "tsd" else
                  {"scale_distance": 1.0 / 0.15 ** 2, "scale_orientation": 0.33} if name == "rnm" else {"params": wl.PDF_PARAMS})
            for k in range(2):
                prep = mt.prepare(M, mM, S, mS, 10, 360, 1000, phi_max, cfg.sensor.angular_res, seed=k, copy=False)
                res = mt.score_prepared(name, prep, **kw)
            t0 = time.perf_counter()
            for k in range(reps):
                prep = mt.prepare(M, mM, S, mS, 10, 360, 1000, phi_max, cfg.sensor.angular_res, seed=10 + k, copy=False)
            tp = (time.perf_counter() - t0) / reps
            t0 = time.perf_counter()
            for k in range(reps):
                prep = mt.prepare(M, mM, S, mS, 10, 360, 1000, phi_max, cfg.sensor.angular_res, seed=10 + k, copy=False)
                res = mt.score_prepared(name, prep, **kw)
            tt = (time.perf_counter() - t0) / reps
            leg[name] = {"prepare_ms": tp * 1e3, "prepare_plus_score_ms": tt * 1e3, "n_hypotheses": int(prep.n_hyp),
                         "best": int(res[-2])}
        leg["what"] = ("match_prepare (1000 trials, 360 control points, counter-based random numbers) + scoring of the resident "
                       "set, from the raw 1081-point model and scene in host memory to the winning transform")
        out["device_prepared"] = leg
    if keep_inputs:
        out["_inputs"] = {"workload": wl, "pose": hs.pose.copy(), "map_scans": scans[:3], "cfg": cfg}
    return out


def large_grid_sweep(device: int = 0, layout_grid: int = 14, rooms=(40.0, 80.0, 160.0, 320.0, 380.0), reps: int = 5,
                     peak_gbs: float | None = None):
    """BASELINE.json configs[2] (SURVEY.md 8d "C3"): pushes into a 16384^2 grid, dense regime (every partition
    allocated), with a 250 m sensor in the middle of an empty square room of growing size, so that the share of
    the map one push rewrites sweeps from ~1 % to most of the field of view.  Free space far from the sensor is
    classified "empty" and streamed by K3 (33x33 cells per partition): the bandwidth regime of the push.
    Returns one row per room: tiles, updates, kernel times (CUDA events inside the library), Gcell-updates/s and
    algorithmic GB/s (32 B per update) of the whole push and of k_update alone."""
    import torch

    from . import capi

    spec = synth.SensorSpec(max_range=250.0)
    cell = 0.025
    g = capi.Grid(cell, 5, layout_grid, device=device)
    g.set_max_truncation(3 * cell)
    g.fill(1.0, 1.0)
    g.set_timing(True)
    side = (1 << layout_grid) * cell
    stream = torch.cuda.ExternalStream(g.stream_ptr, device=torch.device("cuda", device))
    rows = []
    for k, room in enumerate(rooms):
        world = synth.World.room_with_obstacles(side / 2, side / 2, room, room, 0, 1)
        rng = np.random.default_rng(100 + k)
        hs = HostSensor(spec, capi.invert3x3)
        hs.set_scan(synth.scan_from_pose(world, spec, side / 2, side / 2, 0.3, rng))
        hs.T = np.eye(3)
        hs.rays = hs.rays_local.copy()
        hs.ray_norm = 1.0
        hs.transform(synth.pose_matrix(side / 2, side / 2, 0.3))
        sc = hs.scan()
        for _ in range(2):
            g.push(sc)
        g.stage_scan(sc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.sync()
        e0.record(stream)
        for _ in range(reps):
            g.push_staged()
        e1.record(stream)
        g.sync()
        ms = e0.elapsed_time(e1) / reps
        st, km = g.last_push_stats(), g.last_push_kernel_ms()
        upd = st["cell_updates"]
        row = {"room_m": room, "active_tiles": st["active_tiles"], "emptied_tiles": st["emptied_tiles"],
               "tile_share": (st["active_tiles"] + st["emptied_tiles"]) / float(g.n_partitions), "cell_updates": upd,
               "push_ms": ms, "classify_ms": km["classify"], "update_ms": km["update"],
               "gcell_updates_per_s": upd / ms / 1e6, "algorithmic_gbs": 32.0 * upd / ms / 1e6,
               "k_update_algorithmic_gbs": 32.0 * upd / km["update"] / 1e6}
        if peak_gbs:
            row["k_update_frac_of_hbm_peak"] = row["k_update_algorithmic_gbs"] / peak_gbs
        rows.append(row)
    return {"grid": f"{1 << layout_grid}x{1 << layout_grid} @ 2.5 cm, dense, {g.n_partitions} partitions, sensor 270 deg / 250 m",
            "rows": rows}


def raycast_sweep(device: int = 0, main=None, reps: int = 20):
    """BASELINE.json configs[2] asks for a "push + raycast throughput sweep": one 1081-beam RayCastPolar2D cast
    (calcCoordsFromCurrentViewMask, C ABI with host buffers: H2D of the scan + rays, kernel, D2H of points / normals /
    mask inside the timing) on the C1, C2 and C3 maps.  rays/s, march steps/s (fine steps of RayCastPolar2D.cpp:238-247
    + coarse partition skips :225-235, counted by the kernel), algorithmic GB/s at 32 B per fine step.
    main = (name, grid, workload): a map that is already built (the bench's headline grid) is reused."""
    import time

    from . import capi
    rows = []
    for name in ("C1", "C2", "C3"):
        if main is not None and main[0] == name:
            g, wl = main[1], main[2]
            own = False
        else:
            wl = DoubleLaserWorkload(name, invert=capi.invert3x3, n_map=4, n_steps=1)
            cfg = wl.cfg
            g = capi.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid, device=device)
            g.set_max_truncation(cfg.max_truncation)
            for sc in wl.map_scans:
                g.push(sc)
            own = True
        sc, rays = wl.step_scans[0][0], wl.step_rays[0][0]
        for _ in range(3):
            c, nrm, m, cnt = g.raycast_mask(sc, rays)
        t0 = time.perf_counter()
        for _ in range(reps):
            c, nrm, m, cnt = g.raycast_mask(sc, rays)
        dt = (time.perf_counter() - t0) / reps
        fine, coarse = g.raycast_steps()
        rows.append({"map": name, "cells": wl.cfg.cells, "max_range_m": wl.cfg.sensor.max_range, "beams": int(sc.n), "hits": int(cnt),
                     "ms": dt * 1e3, "rays_per_s": sc.n / dt, "fine_steps": int(fine), "coarse_steps": int(coarse),
                     "steps_per_s": (fine + coarse) / dt, "algorithmic_gbs": 32.0 * fine / dt / 1e9,
                     "regime": "dense (bench map)" if not own else "sparse (map built by 8 pushes)"})
        if own:
            del g
    return rows


def secondary_push_benchmark(name: str, device: int = 0, steps: int = 200, peak_gbs: float | None = None):
    """The push leg of bench.py on another configuration (configs[1], the 4096^2 double-laser map, when the headline is
    configs[2]): device-resident value, end-to-end value and the live k_update roofline fraction."""
    import time

    import torch

    from . import capi
    wl = DoubleLaserWorkload(name, invert=capi.invert3x3)
    cfg = wl.cfg
    g = capi.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid, device=device)
    g.set_max_truncation(cfg.max_truncation)
    wl.build_map(g)
    g.set_timing(True)
    stream = torch.cuda.ExternalStream(g.stream_ptr, device=torch.device("cuda", device))
    batches = [capi.ScanBatch(list(st)) for st in wl.step_scans]
    n = len(batches)
    upd = []
    for b in batches:
        g.push_batch(b)
        upd.append(g.last_push_stats()["cell_updates"])
    evs, kms = [], []
    g.sync()
    for i in range(steps):
        g.stage_batch(batches[i % n])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        g.push_staged()
        e1.record(stream)
        evs.append((e0, e1))
        if i % 8 == 0:
            kms.append((g.last_push_kernel_ms()["update"], upd[i % n]))
    g.sync()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    total = sum(upd[i % n] for i in range(steps))
    t0 = time.perf_counter()
    for i in range(steps):
        g.push_batch(batches[i % n])
    e2e_s = time.perf_counter() - t0
    k_ms = float(np.mean([k for k, _ in kms]))
    k_upd = float(np.mean([u for _, u in kms]))
    out = dict(wl.describe(), steps=steps, value_gcell_updates_per_s=total / ms / 1e6, ms_per_step=ms / steps,
               e2e_gcell_updates_per_s=total / e2e_s / 1e9, e2e_ms_per_step=e2e_s / steps * 1e3, k_update_ms=k_ms,
               k_update_algorithmic_gbs=32.0 * k_upd / k_ms / 1e6)
    if peak_gbs:
        out["k_update_frac_of_hbm_peak"] = out["k_update_algorithmic_gbs"] / peak_gbs
    return out
