"""Hypothesis scoring throughput (BASELINE.json configs[3], "C4"): not a pytest test.
  python tests/gpu_perf_match.py [n_hyp]            # one GPU
  torchrun --nproc-per-node N tests/gpu_perf_match.py  # the hypothesis list split over N GPUs, best merged"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ohm_tsd_slam_b200.workload import hypothesis_benchmark

if __name__ == "__main__":
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = hypothesis_benchmark(device=local, n_hyp=int(sys.argv[1]) if len(sys.argv) > 1 else 100000, dist=dist)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out))
    if dist:
        dist.destroy_process_group()
