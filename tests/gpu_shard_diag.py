"""Where a sharded step spends its time (not a pytest test).
torchrun --nproc-per-node 2 tests/gpu_shard_diag.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from ohm_tsd_slam_b200 import capi
from ohm_tsd_slam_b200.sharded import DistBand
from ohm_tsd_slam_b200.workload import MultiRobotWorkload

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = MultiRobotWorkload(world, "C2", invert=capi.invert3x3)
band = DistBand(wl.cell_size, wl.layout_grid, local)
g = band.grid
g.set_max_truncation(wl.max_truncation)
for sc in wl.map_scans:
    band.push(sc)
g.fill(1.0, 1.0, only_uninitialized=True)
band.sync_halos(full=True)
g.set_timing(True)
stream = torch.cuda.ExternalStream(g.stream_ptr, device=torch.device("cuda", local))


def ev():
    return torch.cuda.Event(enable_timing=True)


def run(n, do_sync):
    pe, se = [], []
    g.sync(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        for sc in wl.step_scans[i % len(wl.step_scans)]:
            if not band.stage_and_note(sc):
                continue
            a, b = ev(), ev()
            a.record(stream); g.push_staged(); b.record(stream)
            pe.append((a, b))
        if do_sync:
            a, b = ev(), ev()
            a.record(stream); band.sync_halos(); b.record(stream)
            se.append((a, b))
        else:
            band.dirty_lo.clear(); band.dirty_hi.clear()
    g.sync(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    p = sum(a.elapsed_time(b) for a, b in pe) / n
    s = sum(a.elapsed_time(b) for a, b in se) / n if se else 0.0
    return wall, p, s, len(pe) / n


for do_sync in (False, True, True):
    w, p, s, k = run(100, do_sync)
    print(f"rank {rank} sync={do_sync}: wall {w:.3f} ms/step, pushes {p:.3f} ms ({k:.0f}/step), halo sync {s:.3f} ms", flush=True)
# one push of each kind
for sc in wl.step_scans[0]:
    if band.stage_and_note(sc):
        g.push_staged()
        print(f"rank {rank} push box {g.scan_box(sc)} rows {band.rows[rank]} stats {g.last_push_stats()} {g.last_push_kernel_ms()}", flush=True)
band.sync_halos()
print(f"rank {rank} dirty cols lo {band.dirty_lo.lo, band.dirty_lo.hi} hi {band.dirty_hi.lo, band.dirty_hi.hi}")
dist.destroy_process_group()
