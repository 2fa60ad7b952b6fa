"""world_size-2 gloo tests (CPU) of the host-side logic of the sharded path: band assignment, the min-merge of
per-beam ray crossings with winner payload, and the hypothesis arg-max merge."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ohm_tsd_slam_b200.sharded import (NO_EVENT, DirtyColumns, band_reached, merge_best_hypothesis, merge_first_events, split_rows,
                                       touched_boundaries)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        n = 1081
        # ground truth: per beam an event step and code, owned by exactly one rank; later local events elsewhere
        first_step = rng.integers(0, 3000, n)
        code = rng.choice([0, 1, 2], n, p=[0.8, 0.05, 0.15])
        owner = rng.integers(0, world, n)
        none = rng.uniform(size=n) < 0.1
        keys = np.full(n, NO_EVENT, dtype=np.int64)
        payload = np.zeros((n, 4))
        truth_payload = rng.normal(size=(n, 4))
        for b in range(n):
            if none[b]:
                continue
            if owner[b] == rank:
                keys[b] = 4 * first_step[b] + code[b]
                if code[b] == 0:
                    payload[b] = truth_payload[b]
            elif rng.uniform() < 0.5:  # a later event in this band: must lose
                keys[b] = 4 * (first_step[b] + 1 + rng.integers(0, 50)) + 0
                payload[b] = rng.normal(size=4)
        mask, out, gmin = merge_first_events(torch.from_numpy(keys), torch.from_numpy(payload),
                                             lambda t: dist.all_reduce(t, op=dist.ReduceOp.MIN),
                                             lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM))
        exp_mask = (~none) & (code == 0)
        ok = np.array_equal(mask.numpy(), exp_mask) and np.array_equal(out.numpy()[exp_mask], truth_payload[exp_mask])
        ok = ok and np.all(out.numpy()[~exp_mask] == 0.0)
        # hypothesis arg-max: equal best scores on both ranks -> the lower global index wins
        score, idx = (0.75, 40 + rank * 100) if rank == 0 else (0.75, 7)
        s, i = merge_best_hypothesis(score, idx, lambda t: dist.all_reduce(t, op=dist.ReduceOp.MAX))
        ok = ok and (s == 0.75 and i == 7)
        s, i = merge_best_hypothesis(0.1 * (rank + 1), 5 + rank, lambda t: dist.all_reduce(t, op=dist.ReduceOp.MAX))
        ok = ok and (abs(s - 0.1 * world) < 1e-15 and i == 5 + world - 1)
        s, i = merge_best_hypothesis(0.0, -1, lambda t: dist.all_reduce(t, op=dist.ReduceOp.MAX))
        ok = ok and i == -1
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_merge_over_two_gloo_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_split_rows():
    assert split_rows(128, 8) == [(16 * i, 16 * (i + 1)) for i in range(8)]
    rows = split_rows(32, 5)
    assert rows[0][0] == 0 and rows[-1][1] == 32 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    assert max(e - b for b, e in rows) - min(e - b for b, e in rows) <= 1


def test_boundary_bookkeeping_is_symmetric():
    """Both sides of a band boundary must derive the same dirty partition columns from the (replicated) scans,
    because the halo exchange is a matched send / receive of exactly that slice."""
    rng = np.random.default_rng(11)
    parts = 128
    for world in (2, 3, 8):
        rows = split_rows(parts, world)
        lo = [DirtyColumns() for _ in range(world)]
        hi = [DirtyColumns() for _ in range(world)]
        for _ in range(200):
            x0, y0 = rng.integers(0, parts, 2)
            box = (int(x0), int(y0), int(min(parts - 1, x0 + rng.integers(0, 40))), int(min(parts - 1, y0 + rng.integers(0, 40))))
            reached = [band_reached(box, b, e) for b, e in rows]
            assert any(reached)
            for r, (b, e) in enumerate(rows):
                l, u = touched_boundaries(box, b, e, r, world)
                if l:
                    lo[r].add(box[0], box[2])
                    assert reached[r] and reached[r - 1]
                if u:
                    hi[r].add(box[0], box[2])
                    assert reached[r] and reached[r + 1]
                # a rank that owns rows inside the box is always reached
                if box[1] < e and box[3] >= b:
                    assert reached[r]
        for r in range(world - 1):
            assert (hi[r].lo, hi[r].hi) == (lo[r + 1].lo, lo[r + 1].hi)
        assert not lo[0] and not hi[world - 1]
