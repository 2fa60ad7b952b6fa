"""A TsdGrid sharded in bands of partition rows, one band per GPU (BASELINE.json config 5).

Plumbing only: the kernels live in libtsdslam_b200 (tsdg_create_band & co.); this module moves boundary rows
between bands and merges per-beam first events, with torch.distributed (NCCL over NVLink, or gloo in the CPU
tests of the merge logic) or, for several bands on ONE device, with plain device copies.

  push      every band integrates the (replicated) scan into its own rows            -> tsdg_push_async
            boundary partition rows go to the neighbouring bands (halo)              -> exchange
            replicated borders of the band's top row are refreshed from the halo     -> tsdg_band_push_finish
            boundary rows once more, now with refreshed borders (ray casting reads them)
  raycast   every band marches all beams, evaluates only the steps whose sample it owns -> tsdg_raycast_band_keys
            keys: all-reduce MIN; payload of the winner: mask + all-reduce SUM
"""
from __future__ import annotations

import numpy as np
import torch

from . import capi
from .scan import Scan

NO_EVENT = np.iinfo(np.int64).max  # UINT64_MAX keys never occur below 2^63; int64 max after the cast below


class _Ptr:
    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(ptr: int, count: int, typestr: str, device: int) -> torch.Tensor:
    """torch view of `count` elements of library-owned device memory."""
    return torch.as_tensor(_Ptr(ptr, (count,), typestr), device=torch.device("cuda", device))


def split_rows(parts_y: int, world: int):
    """Partition rows [begin, end) of every band: equal shares, the remainder spread over the first bands."""
    base, rem = divmod(parts_y, world)
    out = []
    b = 0
    for r in range(world):
        e = b + base + (1 if r < rem else 0)
        out.append((b, e))
        b = e
    return out


def merge_first_events(keys: torch.Tensor, payload: torch.Tensor, allreduce_min, allreduce_sum):
    """keys int64[n] (4*step + code, NO_EVENT = none), payload f64[n, 4].  Returns (mask bool[n], payload[n, 4])
    identical on every rank: the earliest event of each beam wins; it is a hit iff its code is 0."""
    gmin = keys.clone()
    allreduce_min(gmin)
    mine = (keys == gmin) & (gmin != NO_EVENT)
    contrib = torch.where(mine[:, None], payload, torch.zeros_like(payload))
    allreduce_sum(contrib)
    mask = (gmin != NO_EVENT) & ((gmin & 3) == 0)
    return mask, contrib, gmin


def merge_best_hypothesis(score: float, index: int, allreduce_max_i64):
    """Arg-max of a score over ranks with first-index tie-break: packs (orderable score bits, ~index)."""
    assert score >= 0.0 or index < 0
    bits = np.float64(max(score, 0.0)).view(np.int64)  # non-negative doubles order like their bit patterns
    # 63 bits of score are too many to share a word with the index: reduce the score first, then the index
    t = torch.tensor([int(bits)], dtype=torch.int64)
    allreduce_max_i64(t)
    best_bits = int(t[0])
    cand = torch.tensor([-(index) if (int(bits) == best_bits and index >= 0) else -NO_EVENT], dtype=torch.int64)
    allreduce_max_i64(cand)  # max of -index = min index
    win = -int(cand[0])
    return (np.int64(best_bits).view(np.float64).item(), win if win != NO_EVENT else -1)


class LocalBands:
    """All bands of a sharded grid on ONE device (tests, and a way to prove that sharding is exact)."""

    def __init__(self, cell_size: float, layout_grid: int, bands: int, device: int = 0):
        self.device = device
        parts_y = (1 << layout_grid) // 32
        self.rows = split_rows(parts_y, bands)
        self.grids = [capi.Grid(cell_size, 5, layout_grid, device=device, band=r) for r in self.rows]
        self.n_partitions = self.grids[0].n_partitions
        self.parts_x = parts_y
        self.dim = 32

    def set_max_truncation(self, v):
        for g in self.grids:
            g.set_max_truncation(v)

    @property
    def bounds(self):
        return self.grids[0].bounds

    def free_footprint(self, *a):
        return all([g.free_footprint(*a) for g in self.grids])

    def _exchange(self):
        for g in self.grids:
            g.sync()
        for i, g in enumerate(self.grids):
            if i > 0:  # my lowest row -> upper halo of the band below
                st, sw, n = g.band_row(0)
                dt, dw, m = self.grids[i - 1].band_row(3)
                assert n == m and n > 0
                device_tensor(dt, n, "<f8", self.device).copy_(device_tensor(st, n, "<f8", self.device))
                device_tensor(dw, n, "<f8", self.device).copy_(device_tensor(sw, n, "<f8", self.device))
            if i + 1 < len(self.grids):  # my highest row -> lower halo of the band above
                st, sw, n = g.band_row(1)
                dt, dw, m = self.grids[i + 1].band_row(2)
                assert n == m and n > 0
                device_tensor(dt, n, "<f8", self.device).copy_(device_tensor(st, n, "<f8", self.device))
                device_tensor(dw, n, "<f8", self.device).copy_(device_tensor(sw, n, "<f8", self.device))
        torch.cuda.synchronize(self.device)

    def push(self, scan: Scan):
        for g in self.grids:
            g.push_async(scan)
        self._exchange()
        for g in self.grids:
            g.band_push_finish()
        self._exchange()

    def last_push_stats(self):
        sts = [g.last_push_stats() for g in self.grids]
        out = dict(sts[0])
        for k in ("cell_updates", "cell_visits", "fallback_cells"):
            out[k] = sum(s[k] for s in sts)
        return out

    def partition_states(self):
        return self.grids[0].partition_states()  # replicated

    def download_partition(self, p: int):
        py = p // self.parts_x
        for (b, e), g in zip(self.rows, self.grids):
            if b <= py < e:
                return g.download_partition(p)
        return None

    def raycast_mask(self, scan: Scan, rays_world, coords=None, normals=None):
        n = scan.n
        ks, ps = [], []
        for g in self.grids:
            kp, pp = g.raycast_band_keys(scan, rays_world)
            ks.append(device_tensor(kp, n, "<i8", self.device).clone())
            ps.append(device_tensor(pp, 4 * n, "<f8", self.device).clone().view(n, 4))
        K = torch.stack(ks)

        def amin(t):
            t.copy_(K.min(dim=0).values)

        results = []
        for r in range(len(self.grids)):
            # what rank r would compute; the sum runs over every rank's masked contribution
            def asum(t, r=r):
                g = K.min(dim=0).values
                tot = torch.zeros_like(t)
                for q in range(len(self.grids)):
                    mine = (K[q] == g) & (g != NO_EVENT)
                    tot += torch.where(mine[:, None], ps[q], torch.zeros_like(ps[q]))
                t.copy_(tot)

            results.append(merge_first_events(ks[r], ps[r], amin, asum))
        mask, payload, gmin = results[0]
        for m2, p2, _ in results[1:]:
            assert torch.equal(mask, m2) and torch.equal(payload, p2)
        mask = mask.cpu().numpy().astype(np.uint8)
        payload = payload.cpu().numpy()
        coords = np.zeros((n, 2)) if coords is None else coords
        normals = np.zeros((n, 2)) if normals is None else normals
        coords[mask > 0] = payload[mask > 0, :2]
        normals[mask > 0] = payload[mask > 0, 2:]
        return coords, normals, mask, int(mask.sum())


class DistBand:
    """One band per rank over torch.distributed (backend nccl on GPUs)."""

    def __init__(self, cell_size: float, layout_grid: int, device: int):
        import torch.distributed as dist
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.device = device
        parts_y = (1 << layout_grid) // 32
        self.rows = split_rows(parts_y, self.world)
        self.grid = capi.Grid(cell_size, 5, layout_grid, device=device, band=self.rows[self.rank])
        self._views = None

    def _row_views(self):
        if self._views is None:
            v = {}
            for which in range(4):
                t, w, n = self.grid.band_row(which)
                v[which] = None if n == 0 else (device_tensor(t, n, "<f8", self.device), device_tensor(w, n, "<f8", self.device))
            self._views = v
        return self._views

    def exchange(self):
        """Boundary rows to the neighbouring bands, stream-ordered (no host synchronisation): the collective
        stream waits for the library's stream, the library's stream waits for the collective."""
        dist = self.dist
        v = self._row_views()
        cur = torch.cuda.current_stream(self.device)
        self.grid.stream_order(cur.cuda_stream, 0)
        ops = []
        if self.rank > 0:  # lowest row down, lower halo from below
            for k in (0, 1):
                ops.append(dist.P2POp(dist.isend, v[0][k], self.rank - 1))
                ops.append(dist.P2POp(dist.irecv, v[2][k], self.rank - 1))
        if self.rank + 1 < self.world:
            for k in (0, 1):
                ops.append(dist.P2POp(dist.isend, v[1][k], self.rank + 1))
                ops.append(dist.P2POp(dist.irecv, v[3][k], self.rank + 1))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()  # NCCL: orders the current stream after the transfer, does not block the host
        self.grid.stream_order(cur.cuda_stream, 1)
        self.halo_dirty = False

    def _push_tail(self):
        # phase 2 needs the upper neighbour's fresh first row; the refreshed borders reach the neighbours'
        # halos lazily, before the next ray cast (nothing else reads a halo's border cells)
        self.exchange()
        self.grid.band_push_finish()
        self.halo_dirty = True

    def push(self, scan: Scan):
        self.grid.push_async(scan)
        self._push_tail()

    def push_staged(self):
        self.grid.push_staged()
        self._push_tail()

    def raycast_mask(self, scan: Scan, rays_world):
        dist = self.dist
        n = scan.n
        if getattr(self, "halo_dirty", False):
            self.exchange()
        kp, pp = self.grid.raycast_band_keys(scan, rays_world)
        keys = device_tensor(kp, n, "<i8", self.device)
        payload = device_tensor(pp, 4 * n, "<f8", self.device).view(n, 4)
        mask, out, _ = merge_first_events(keys, payload, lambda t: dist.all_reduce(t, op=dist.ReduceOp.MIN),
                                          lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM))
        mask = mask.cpu().numpy().astype(np.uint8)
        out = out.cpu().numpy()
        return out[:, :2], out[:, 2:], mask, int(mask.sum())
