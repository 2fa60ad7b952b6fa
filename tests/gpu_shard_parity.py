"""A grid sharded over several PROCESSES (one band per rank, halo rows over CUDA IPC peer stores or NCCL) equals the
unsharded grid bit for bit: every partition after pushes, every ray cast (min-merge of the bands' crossings).
  torchrun --nproc-per-node 2 tests/gpu_shard_parity.py [peer|nccl]             # one GPU per rank, NCCL plumbing
  torchrun --nproc-per-node 2 tests/gpu_shard_parity.py peer --one-device      # all ranks on GPU 0, gloo plumbing:
                                                                                  # what tests/test_sharded_multiproc.py runs
                                                                                  # under pytest -m gpu on a one-GPU box"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from ohm_tsd_slam_b200 import capi, synth
from ohm_tsd_slam_b200.scan import HostSensor
from ohm_tsd_slam_b200.sharded import DistBand
from tests.harness import same

transport = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "peer"
one_device = "--one-device" in sys.argv
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = 0 if one_device else int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
if one_device:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = synth.config("C1")
whole = capi.Grid(cfg.cell_size, 5, cfg.layout_grid, device=local)      # every rank keeps its own unsharded copy
band = DistBand(cfg.cell_size, cfg.layout_grid, local, transport=transport)
whole.set_max_truncation(cfg.max_truncation)
band.grid.set_max_truncation(cfg.max_truncation)
hs = HostSensor(cfg.sensor, capi.invert3x3)
parts_x = (1 << cfg.layout_grid) // 32
b, e = band.rows[rank]
bad = 0
checked = 0
for k, (pose, r) in enumerate(cfg.scans(10)):
    hs.set_scan(r)
    hs.T = synth.pose_matrix(*pose)
    sc = hs.scan()
    if k > 0:
        rays = hs.normalized_rays(cfg.cell_size).copy()
        c1, n1, m1, k1 = whole.raycast_mask(sc, rays)
        c2, n2, m2, k2 = band.raycast_mask(sc, rays)       # halo sync + flags merge + min-merge of the crossings
        if not (same(m1, m2) and same(c1[m1 > 0], c2[m2 > 0]) and same(n1[m1 > 0], n2[m2 > 0]) and k1 == k2 and k1 > 0):
            bad += 1
            print(f"rank {rank} scan {k}: raycast differs ({k1} vs {k2} hits)", flush=True)
    whole.push(sc)
    band.push(sc)
    if k % 3 == 2:  # several pushes per synchronisation
        band.sync_halos()
        band.grid.sync()
        st, _ = whole.partition_states()
        for p in np.nonzero(st == 2)[0]:
            if not (b <= p // parts_x < e):
                continue
            ta, wa = whole.download_partition(int(p))
            got = band.grid.download_partition(int(p))
            checked += 1
            if got is None or not (same(ta, got[0]) and same(wa, got[1])):
                bad += 1
                if bad < 4:
                    print(f"rank {rank} scan {k}: partition {p} differs", flush=True)
t = torch.tensor([bad, checked], device="cpu" if one_device else "cuda")
dist.all_reduce(t)
if rank == 0:
    print(f"transport {transport}, {world} GPUs: {int(t[1])} partition comparisons, {int(t[0])} mismatches -> {'OK' if int(t[0]) == 0 else 'FAIL'}")
dist.destroy_process_group()
sys.exit(0 if int(t[0]) == 0 else 1)
