#!/usr/bin/env python
"""bench.py -- TsdGrid::push throughput (+ raycast + ICP) of the B200-native hot path.  N = 1: BASELINE.json
configs[2], the largest single-GPU configuration (16384^2 map, dense regime, the double-laser robot of configs[1] in a
360 x 270 m hall with 1024 obstacles: addTsd-dominated); configs[1] (4096^2) rides along as `c2_double_laser`.
N > 1: N robots on one (8192 N)^2 grid sharded in bands (configs[4]).  One JSON line on stdout (rank 0).

  python bench.py --gpus 1 --steps 50 --warmup 5            # CUDA arm (the product, through its C ABI)
  python bench.py --impl reference --steps 3 --warmup 1     # the reference's own CPU code (oracle/_ref)
  torchrun ... bench.py --gpus N ...                        # N ranks, one replica of the workload per GPU

A step = one scan cycle of both lasers: two TsdGrid::push calls.  `value` times the pushes with the scans
already staged in HBM; `e2e` times the same pushes through tsdg_push() with host buffers (H2D of the scan and
D2H of the push statistics inside the timed region)."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tsd_push_cell_updates_per_s"
UNIT = "Gcell-updates/s"
ALG_BYTES_PER_UPDATE = 32  # read + write of {tsd, weight} in FP64


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def run_cuda(args):
    import torch
    import torch.distributed as dist

    from ohm_tsd_slam_b200 import capi
    from ohm_tsd_slam_b200.workload import DoubleLaserWorkload

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if capi.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1:
        wl = DoubleLaserWorkload(args.workload, invert=capi.invert3x3)
        cfg = wl.cfg
        grid = capi.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid, device=local)
        band = None
    else:
        # N robots of the same kind on one grid sharded in N bands of partition rows, one band per GPU (weak
        # scaling; BASELINE.json configs[4]).  Every rank sees every scan, integrates those that can reach its band
        # (tsdg_scan_box) into its own rows -- no communication -- and synchronises the boundary rows with its
        # neighbours once per step (P2P over NCCL), as a SLAM cycle that ray-casts after pushing would.
        from ohm_tsd_slam_b200.sharded import DistBand
        from ohm_tsd_slam_b200.workload import MultiRobotWorkload, secondary_push_benchmark
        # The weak-scaling unit is one robot of this kind on one GPU.  The default N = 1 run of this script headlines the
        # 16384^2 map instead (BASELINE configs[2]), so the N > 1 line carries its own reference: rank 0 first runs one
        # robot alone on its 4096^2 grid (the others wait).
        one_gpu = None
        if rank == 0:
            one_gpu = secondary_push_benchmark(args.workload, device=local, steps=min(args.steps, 200), peak_gbs=measured_peak()[0])
        dist.barrier()
        wl = MultiRobotWorkload(world, args.workload, invert=capi.invert3x3)
        cfg = wl.cfg
        band = DistBand(wl.cell_size, wl.layout_grid, local)
        grid = band.grid
    grid.set_max_truncation(cfg.max_truncation)

    for sc in wl.map_scans:
        band.push(sc) if band else grid.push(sc)
    grid.fill(1.0, 1.0, only_uninitialized=True)
    if band:
        band.sync_halos(full=True)
    grid.set_timing(True)
    launches0 = capi.kernel_launches()
    stream = torch.cuda.ExternalStream(grid.stream_ptr, device=torch.device("cuda", local))
    n_steps = len(wl.step_scans)

    def ev_pair():
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---------------- device-resident leg: `value` times the pushes (and the halo synchronisation of a sharded grid)
    # with CUDA events on the library's stream; the H2D staging of each scan (tsdg_stage_scan, from pinned host
    # memory) sits between the event pairs, outside them.
    # A step = one scan cycle: the two lasers of every robot.  The two scans of a robot form one batch
    # (tsdg_stage_batch / tsdg_push_batch: the result of two TsdGrid::push calls, one classification + one update
    # launch for both; --no-batch pushes them one by one).
    def make_batches(sc):
        if band:  # only what reaches this rank's band; up to four scans (two robots) per launch pair
            b, e = band.rows[band.rank]
            from ohm_tsd_slam_b200.sharded import band_reached
            sc = [s1 for s1 in sc if band_reached(band._box(s1), b, e)]
        if args.no_batch:
            groups = [[s1] for s1 in sc]
        else:
            groups, k = [], 0
            while k < len(sc):
                take = 4 if len(sc) - k >= 4 else (2 if len(sc) - k >= 2 else 1)
                groups.append(list(sc[k:k + take]))
                k += take
        return [capi.ScanBatch(gr) for gr in groups]  # the tsd_scan_t arrays are laid out once

    step_batches = [make_batches(sc) for sc in wl.step_scans]

    def batches(i):
        return step_batches[i % n_steps]

    def note_step(i):
        # every rank looks at every scan of the step: which band boundaries it dirties (replicated book-keeping)
        for s1 in wl.step_scans[i % n_steps]:
            band.note_scan(band._box(s1))
        band.flags_dirty = True

    def resident_step(i, acc, samples=None):
        if band:
            note_step(i)
        for b in batches(i):
            grid.stage_batch(b)
            e0, e1 = ev_pair()
            e0.record(stream)
            grid.push_staged()
            e1.record(stream)
            acc.append((e0, e1))
            if samples is not None:  # (synchronises, between the event pairs) kernel times + update count of this launch
                samples.append((grid.last_push_kernel_ms(), grid.last_push_stats()["cell_updates"]))
        # Pushes need no communication (each band integrates the scans that reach it); the halo rows are synchronised
        # before something READS across a band boundary -- the sharded ray cast below does, and its time includes it.
        # --halo-every-step puts one synchronisation into every step, as round 1 did.
        if band and args.halo_every_step:
            e0, e1 = ev_pair()
            e0.record(stream)
            band.sync_halos()
            e1.record(stream)
            acc.append((e0, e1))

    for i in range(args.warmup):
        resident_step(i, [])
    grid.sync()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = []
    kms = []
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        resident_step(i, evs, kms if i % 8 == 0 else None)  # per-kernel times: sampled, not every step
    grid.sync()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)

    # exact update count of one step on this rank, and of its heaviest push (the roofline launch)
    def host_step(i, count=None):
        if band:
            note_step(i)
        for b in batches(i):
            grid.push_batch(b)  # blocking: H2D of the scans, kernels, D2H of the statistics, synchronise
            if count is not None:
                count.append(grid.last_push_stats()["cell_updates"])
                scans_pushed[0] += len(b)
        if band and args.halo_every_step:
            band.sync_halos()
            grid.sync()

    upd_steps = []
    scans_pushed = [0]
    for i in range(n_steps):
        c = []
        host_step(i, c)
        upd_steps.append(c)
    pushes_per_step = len(upd_steps[0])
    scans_per_step = scans_pushed[0] / n_steps  # scans this rank integrates per step
    upd_per_step = float(np.mean([sum(c) for c in upd_steps]))
    upd_total_dev = sum(sum(upd_steps[i % n_steps]) for i in range(args.steps))
    # roofline samples: full-size pushes only (on a sharded grid a rank also sees the tail of its neighbour's scans)
    big = [(k, u) for k, u in kms if u > 0.25 * max(u2 for _, u2 in kms)]
    upd_avg_per_push = float(np.mean([u for _, u in big]))
    kms = [k for k, _ in big]

    # ---------------- end-to-end leg: host buffers through tsdg_push (blocking; H2D + statistics D2H inside)
    for i in range(args.warmup):
        host_step(i)
    barrier()
    e2e_updates = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        c = []
        host_step(i, c)
        e2e_updates += sum(c)
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # ---------------- one halo synchronisation after all those pushes (what a read across the band boundary waits for)
    halo_ms = None
    if band:
        barrier()
        e0, e1 = ev_pair()
        e0.record(stream)
        band.sync_halos()
        e1.record(stream)
        grid.sync()
        halo_ms = e0.elapsed_time(e1)

    # ---------------- raycast + ICP through the C ABI (host buffers), for the scans/s part of the metric
    icp = capi.Icp(30, 0.4, 0.02, grid.bounds, device=local)
    sc0, rays0 = wl.step_scans[0][0], wl.step_rays[0][0]
    hs = wl.sensors[0]
    scene = None
    rc_ms = icp_ms = loc_ms = None
    loc_out = None
    reps = max(3, min(args.steps, 20))
    def raycast():
        return band.raycast_mask(sc0, rays0) if band else grid.raycast_mask(sc0, rays0)

    for _ in range(2):
        c, nrm, m, cnt = raycast()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        c, nrm, m, cnt = raycast()
    rc_ms = (time.perf_counter() - t0) / reps * 1e3
    valid = (~np.isinf(sc0.ranges)) & (sc0.mask != 0)
    scene = np.stack([hs.rays_local[0, valid] * sc0.ranges[valid], hs.rays_local[1, valid] * sc0.ranges[valid]], axis=1)
    icp_out = None
    if cnt > 2:
        for _ in range(2):
            icp_out = icp.run(c[m > 0], nrm[m > 0], scene, sc0.pose)
        t0 = time.perf_counter()
        for _ in range(reps):
            icp_out = icp.run(c[m > 0], nrm[m > 0], scene, sc0.pose)
        icp_ms = (time.perf_counter() - t0) / reps * 1e3
        if not band:
            # the same step through the fused entry point: the model never leaves the device
            for _ in range(2):
                loc_out = icp.localize(grid, sc0, rays0, scene)
            t0 = time.perf_counter()
            for _ in range(reps):
                loc_out = icp.localize(grid, sc0, rays0, scene)
            loc_ms = (time.perf_counter() - t0) / reps * 1e3

    # ---------------- map publication: RayCastAxisAligned2D::calcCoords + grid2ColorImage on the device,
    # host buffers in and out (occupancy grid cells_x * cells_y bytes both ways, crossings, 1024^2 RGB image)
    pub = None
    if world == 1:
        occ = np.full(grid.cells * grid.cells, -1, dtype=np.int8)
        grid.axis_map(occupied=occ)
        grid.color_image(1024, 1024)
        t0 = time.perf_counter()
        for _ in range(3):
            pc, _, occ = grid.axis_map(occupied=occ)
        t_axis = (time.perf_counter() - t0) / 3
        t0 = time.perf_counter()
        for _ in range(3):
            grid.color_image(1024, 1024)
        t_img = (time.perf_counter() - t0) / 3
        pub = {"axis_aligned_map_ms": t_axis * 1e3, "crossings": int(len(pc)), "free_cells": int((occ == 0).sum()),
               "color_image_1024_ms": t_img * 1e3}

    # ---------------- hypothesis scoring (BASELINE.json configs[3]: 10^5 hypotheses per scan; split over the ranks)
    from ohm_tsd_slam_b200.workload import hypothesis_benchmark
    want_cpu = not (args.no_cpu_baseline or world > 1)
    hyp = hypothesis_benchmark(device=local, n_hyp=100000, reps=20, dist=dist if world > 1 else None, keep_inputs=want_cpu)
    hyp_inputs = hyp.pop("_inputs", None)

    # ---------------- ray-cast sweep (BASELINE.json configs[2] "push + raycast throughput sweep"): one 1081-beam cast per
    # map size through the C ABI (host buffers), rays/s and march steps/s; L2 hit rates come from the ncu capture
    c2 = None
    rc_sweep = None
    if world == 1 and not args.no_sweep:
        from ohm_tsd_slam_b200.workload import raycast_sweep, secondary_push_benchmark
        rc_sweep = raycast_sweep(device=local, main=(args.workload, grid, wl))
        if args.workload != "C2":
            c2 = secondary_push_benchmark("C2", device=local, steps=min(args.steps, 200), peak_gbs=measured_peak()[0])

    # ---------------- large-map push sweep (BASELINE.json configs[2]): the bandwidth regime of the push
    sweep = None
    if world == 1 and not args.no_sweep:
        from ohm_tsd_slam_b200.workload import large_grid_sweep
        sweep = large_grid_sweep(device=local, layout_grid=14, peak_gbs=measured_peak()[0])

    # ---------------- K2-only / K3-only: the same launches with one kind of work switched off in the update kernel
    # (tsdg_set_update_filter): which part of the roofline fraction is addTsd (K2) and
    # which is increaseEmptiness streaming (K3)
    split = None
    if world == 1:
        split = {}
        for name, mask in (("k2_addTsd_only", 2), ("k3_increaseEmptiness_only", 1)):
            rows = []
            for i in range(n_steps):
                for b in batches(i):
                    grid.set_update_filter(mask)
                    grid.push_batch(b)
                    upd = grid.last_push_stats()["cell_updates"]
                    grid.stage_batch(b)
                    ts = []
                    for _ in range(5):
                        grid.push_staged()
                        ts.append(grid.last_push_kernel_ms()["update"])
                    rows.append((upd, float(np.median(ts))))
            grid.set_update_filter(0)
            u = float(np.mean([r[0] for r in rows]))
            t = float(np.mean([r[1] for r in rows]))
            split[name] = {"cell_updates_per_launch": u, "k_update_ms": t, "algorithmic_gbs": ALG_BYTES_PER_UPDATE * u / t / 1e6,
                           "frac_of_hbm_peak": ALG_BYTES_PER_UPDATE * u / t / 1e6 / measured_peak()[0]}
        # (the filtered pushes leave a map the reference would not have computed: this leg runs last on this grid)

    launches = capi.kernel_launches() - launches0

    # max over ranks of the timed durations, sum of the work
    dev_ms_t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    work_t = torch.tensor([float(upd_total_dev), float(e2e_updates), upd_per_step], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dev_ms_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(work_t, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max = dev_ms_t.tolist()
    work_dev, work_e2e, upd_per_step_all = work_t.tolist()

    if rank == 0:
        peak, peak_src = measured_peak()
        # DRAM traffic of the roofline kernel: only from an ncu capture of THIS workload (profiles/rNN_metrics.json keys
        # "k_update@<workload>"); none for the sharded runs
        traffic, traffic_src = profiled_traffic(f"k_update@{args.workload}") if world == 1 else (None, None)
        upd_ms = float(np.mean([k["update"] for k in kms]))
        achieved = ALG_BYTES_PER_UPDATE * upd_avg_per_push / (upd_ms * 1e-3) / 1e9
        value = work_dev / (dev_ms_max * 1e-3) / 1e9
        e2e_value = work_e2e / (e2e_ms_max * 1e-3) / 1e9
        n = sc0.n
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "cuda",
            "config": dict(wl.describe(), parallelism=(f"one band of partition rows per GPU ({world} bands), scans replicated, pushes without "
                                                       f"communication; before a read across a band boundary (the sharded ray cast) boundary rows are stored "
                                                       f"into the neighbours' halo rows by one kernel over CUDA-IPC peer mappings (NVLink P2P stores), and "
                                                       f"the bands' ray crossings are min-merged by the marching kernel's own P2P stores + a merge kernel"
                                                       if world > 1 else "single GPU"),
                           cell_updates_per_step=upd_per_step_all, push_launch_pairs_per_step_rank0=pushes_per_step,
                           batched=not args.no_batch),
            "hbm_gbs_algorithmic": value * ALG_BYTES_PER_UPDATE,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(scans_per_step * (n * 8 + n + 8 * 25)), "d2h_bytes_per_step": pushes_per_step * 24 * 4,
                    "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_update (TsdGrid::push cell update, K2+K3)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel_ms": upd_ms, "algorithmic_bytes_per_launch": ALG_BYTES_PER_UPDATE * upd_avg_per_push,
                         "split": split},
            "push_kernel_ms": {k: float(np.mean([x[k] for x in kms])) for k in kms[0]},
            "raycast_icp": {"raycast_ms": rc_ms, "icp_ms": icp_ms, "raycast_hits": int(cnt),
                            "localize_ms": loc_ms,
                            "localize": ({"pairs": int(loc_out[2]), "iterations": int(loc_out[3]), "n_model": int(loc_out[5]),
                                          "what": "tsdg_localize: ray cast + maskMatrix + Icp::iterate, model kept on the device, host buffers in and out"}
                                         if loc_out else None),
                            "scans_per_s": (1e3 / (rc_ms + icp_ms)) if icp_ms else None,
                            "scan_ms_push_localize": (loc_ms + e2e_ms_max / args.steps / max(scans_per_step, 1)) if loc_ms else None,
                            "scan_ms_push_raycast_icp": (rc_ms + icp_ms + e2e_ms_max / args.steps / max(scans_per_step, 1)) if icp_ms else None,
                            "icp": None if icp_out is None else {"pairs": icp_out[2], "iterations": icp_out[3]}},
            "one_gpu_same_workload": (None if world == 1 else {k: one_gpu[k] for k in ("workload", "value_gcell_updates_per_s", "ms_per_step",
                                                                                       "e2e_gcell_updates_per_s", "k_update_frac_of_hbm_peak")}),
            "halo_sync_ms_rank0": halo_ms,
            "halo_every_step": bool(args.halo_every_step),
            "map_publication": pub,
            "hypothesis_scoring": hyp,
            "raycast_sweep": rc_sweep,
            "c2_double_laser": c2,
            "k3_streaming_sweep": sweep,
            "clocks": sampler.summary(),
            "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            b = cpu_baseline(args.workload, steps=100000, warmup=1, threads=None, budget_s=12.0)  # ~12 s of CPU work
            line["cpu_baseline"] = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}
            if pub is not None and "map_publication_ms" in b:
                pub["reference_cpu_ms"] = b["map_publication_ms"]
            if hyp_inputs is not None:
                hyp["cpu_port_hypotheses_per_s"] = cpu_hypothesis_baseline(hyp_inputs, sample=1000)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_hypothesis_baseline(inp, sample: int):
    """The three scorers of the plain-C port (one core) on the first `sample` hypotheses of the same workload."""
    from ohm_tsd_slam_b200 import synth
    from ohm_tsd_slam_b200.scan import HostSensor
    from oracle import port
    wl, cfg = inp["workload"], inp["cfg"]
    gp = port.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
    gp.set_max_truncation(cfg.max_truncation)
    hp = HostSensor(cfg.sensor, port.invert3x3)
    for pose, r in inp["map_scans"]:
        hp.set_scan(r)
        hp.T = np.eye(3)
        hp.rays = hp.rays_local.copy()
        hp.ray_norm = 1.0
        hp.transform(synth.pose_matrix(*pose))
        gp.push(hp.scan())
    h = wl.hyps[:sample]
    cpu = {}
    t0 = time.perf_counter()
    port.score_tsd(gp, h, wl.M, wl.S, wl.phi_m, wl.phi_s, wl.phi_max, wl.control, inp["pose"], 0.25)
    cpu["tsd"] = sample / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    port.score_rnm(h, wl.M, wl.S, wl.phi_m, wl.phi_s, wl.phi_max, wl.control, wl.phi_control, wl.model_valid, wl.phi_valid,
                   wl.theta_min, wl.theta_max, 1.0 / 0.15 ** 2, 0.33, wl.control.shape[1] // 3)
    cpu["rnm"] = sample / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    port.score_pdf(h, wl.M, wl.S, wl.phi_m, wl.phi_s, wl.phi_max, wl.control, wl.model_angles, wl.model_dists, wl.PDF_PARAMS)
    cpu["pdf"] = sample / (time.perf_counter() - t0)
    return dict(cpu, sample=sample, cores=1)


def profiled_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the newest committed
    `ncu --set full` summary (profiles/rNN_metrics.json, written by profiles/summarize.py); (None, None) if absent."""
    import glob
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*_metrics.json")))
    for f in reversed(files):
        try:
            m = json.load(open(f)).get(kernel)
            if m:
                return m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"], os.path.join("profiles", os.path.basename(f))
        except (OSError, ValueError, KeyError):
            continue
    return None, None


def cpu_baseline(workload: str, steps: int, warmup: int, threads, budget_s=None):
    """The reference's own push (oracle/_ref, compiled from the reference sources) or, where that prebuilt
    library is absent, the plain-C port, on the host cores; bounded sample of the same workload."""
    from ohm_tsd_slam_b200.workload import DoubleLaserWorkload
    from oracle import port, ref
    use_ref = ref.available()
    inv = ref.invert if use_ref else port.invert3x3
    wl = DoubleLaserWorkload(workload, invert=inv, n_map=2)
    cfg = wl.cfg
    cores = os.cpu_count() or 1
    if use_ref:
        nthreads = threads or cores
        ref.set_threads(nthreads)
        g = ref.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
        g.set_max_truncation(cfg.max_truncation)
        sensor = ref.Sensor(cfg.sensor)

        def push(sc):
            sensor.set_data(sc.ranges, sc.mask)
            sensor.pose = sc.pose
            g.push(sensor)

        kind = "reference"
    else:
        nthreads = 1
        g = port.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
        g.set_max_truncation(cfg.max_truncation)
        push = g.push
        kind = "port"
    # Cell updates per push are a function of the scan and of the allocation state only (dense regime: everything
    # allocated), so they repeat with the scan cycle: count them once per scan with the port (the reference keeps
    # no such statistic), on a second grid that sees the same pushes.
    pg = port.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
    pg.set_max_truncation(cfg.max_truncation)
    for sc in wl.map_scans:
        push(sc)
        pg.push(sc)
    g.fill(1.0, 1.0, only_uninitialized=True)
    pg.fill(1.0, 1.0, only_uninitialized=True)
    n_steps = len(wl.step_scans)
    counts = []
    for i in range(n_steps):
        c = []
        for sc in wl.step_scans[i]:
            pg.push(sc)
            c.append(pg.last_push_stats()["cell_updates"])
        counts.append(c)
    for i in range(max(warmup, 1)):
        for sc in wl.step_scans[i % n_steps]:
            push(sc)
    updates = 0
    t = 0.0
    done = 0
    for i in range(steps):
        for k, sc in enumerate(wl.step_scans[i % n_steps]):
            t0 = time.perf_counter()
            push(sc)
            t += time.perf_counter() - t0
            updates += counts[i % n_steps][k]
        done += 1
        if budget_s is not None and t >= budget_s and done >= 3:
            break
    extra = {}
    if use_ref:  # the map publisher's two calls on the same map (serial code in the reference)
        t0 = time.perf_counter()
        g.axis_map()
        t1 = time.perf_counter()
        g.color_image(1024, 1024)
        extra["map_publication_ms"] = {"axis_aligned_map": (t1 - t0) * 1e3, "color_image_1024": (time.perf_counter() - t1) * 1e3}
    return {**extra, "value": updates / t / 1e9, "unit": UNIT, "cores": nthreads, "kind": kind, "steps": done,
            "sample": f"{2 * done} TsdGrid::push calls ({done} step(s)) of the same workload after the same map build, "
                      f"{updates} cell updates, {t:.2f} s on {nthreads} thread(s) of {cores} host cores",
            "ms_per_step": t / done * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # a step = one scan cycle of both lasers (two pushes, ~35 ms on 16 cores): K steps as asked, capped so that the
    # run ends within a few minutes
    steps = max(1, min(args.steps, 5000))
    warmup = max(0, min(args.warmup, 20))
    b = cpu_baseline(args.workload, steps=steps, warmup=warmup, threads=None, budget_s=150.0)
    steps = b["steps"]  # (fewer than asked only if 150 s of pushes were not enough)
    from ohm_tsd_slam_b200.workload import DoubleLaserWorkload, MultiRobotWorkload
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # the CUDA arm's workload at N GPUs is N robots of this kind on one sharded grid; the reference has one grid in
        # host memory and one process: its push rate per robot is what it is for one robot (sampled: robot 0)
        wl = MultiRobotWorkload(world, args.workload, n_map=0, n_steps=1, invert=lambda T: np.eye(3))
        b["sample"] += f"; robot 0 of the {world}-robot workload on its own {1 << wl.cfg.layout_grid}^2 grid"
    else:
        wl = DoubleLaserWorkload(args.workload, invert=lambda T: np.eye(3), n_map=0, n_steps=1)
    line = {
        "metric": METRIC, "value": b["value"], "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps,
        "warmup": warmup, "ms_per_step": b["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": dict(wl.describe(), parallelism="host cores (OpenMP)"),
        "cpu_baseline": {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": b["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default=None, help="C3 (default at 1 GPU) or C2; N > 1 always uses C2 robots")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch", action="store_true", help="push the two lasers of a robot one by one (two launch pairs)")
    ap.add_argument("--halo-every-step", action="store_true", help="N > 1: synchronise the halo rows inside every timed step")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 16384^2 push sweep (4.6 GB of HBM, a few seconds)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.workload is None:
        args.workload = "C3" if int(os.environ.get("WORLD_SIZE", "1")) == 1 and args.gpus <= 1 else "C2"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
