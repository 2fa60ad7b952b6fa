"""The multi-PROCESS sharded path (what bench.py times at N > 1: one band per rank, halo rows stored into the
neighbour's memory by k_halo_sync over CUDA IPC mappings, ray crossings min-merged over the ranks) under the GPU test
tier, which has one GPU: two ranks share device 0 (CUDA IPC works between processes on one device), gloo does the
plumbing.  tests/gpu_shard_parity.py checks every owned partition and every ray cast against an unsharded grid, bit
for bit."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_bands_in_separate_processes_equal_the_unsharded_grid(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "gpu_shard_parity.py"), "peer", "--one-device"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "0 mismatches -> OK" in out.stdout, out.stdout[-2000:]
    # the comparison really happened: hundreds of partitions, on every rank
    n = int(out.stdout.split(" partition comparisons")[0].split()[-1])
    assert n > 300, out.stdout[-2000:]
