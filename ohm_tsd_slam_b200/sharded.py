"""A TsdGrid sharded in bands of partition rows, one band per GPU (BASELINE.json config 5).

Plumbing only: the kernels live in libtsdslam_b200 (tsdg_create_band & co.); this module moves boundary rows
between bands and merges per-beam first events, with torch.distributed (NCCL over NVLink, or gloo in the CPU
tests of the merge logic) or, for several bands on ONE device, with plain device copies.

  push      every band a scan can reach (tsdg_scan_box) integrates it into its own rows -> tsdg_push_async
            no communication: any number of pushes, from any number of sensors
  sync      before the map is READ across a band boundary (ray casting):
            boundary partition rows go to the neighbouring bands (halo)              -> exchange
            the band's top row takes its top/corner border strips from the halo      -> tsdg_band_push_finish
            boundary rows once more, now with refreshed borders
            (only the partition columns pushed to since the last sync are sent)
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


def band_reached(box, row_begin: int, row_end: int) -> bool:
    """Does a scan with partition box (px0, py0, px1, py1) reach the band [row_begin, row_end)?  The band keeps a
    replica of the allocation state of the rows next to it (halo), so those count as reached too."""
    return box[1] <= row_end and box[3] >= row_begin - 1


def touched_boundaries(box, row_begin: int, row_end: int, rank: int, world: int):
    """(lower, upper): does a scan with this partition box dirty the boundary below / above the band?  A boundary
    between rows e-1 | e is dirty when the box contains either row; both neighbours evaluate the same test."""
    lower = rank > 0 and box[1] <= row_begin and box[3] >= row_begin - 1
    upper = rank + 1 < world and box[1] <= row_end and box[3] >= row_end - 1
    return lower, upper


class DirtyColumns:
    """Partition columns of a band boundary that were pushed to since the last halo synchronisation.  Every rank
    sees every scan, so both sides of a boundary compute the same range without talking to each other."""

    def __init__(self):
        self.lo, self.hi = None, None

    def add(self, px0: int, px1: int):
        self.lo = px0 if self.lo is None else min(self.lo, px0)
        self.hi = px1 if self.hi is None else max(self.hi, px1)

    def clear(self):
        self.lo, self.hi = None, None

    def __bool__(self):
        return self.lo is not None


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

    def __init__(self, cell_size: float, layout_grid: int, bands: int, device: int = 0, peer: bool = False):
        """peer: synchronise halos with the one-kernel peer-memory path (tsdg_band_halo_sync) instead of device
        copies + tsdg_band_push_finish."""
        self.device = device
        self.peer = peer
        parts_y = (1 << layout_grid) // 32
        self.rows = split_rows(parts_y, bands)
        self.grids = [capi.Grid(cell_size, 5, layout_grid, device=device, band=r) for r in self.rows]
        self.n_partitions = self.grids[0].n_partitions
        self.parts_x = parts_y
        self.dim = 32
        self.dirty_cols = DirtyColumns()
        if peer:
            for i, g in enumerate(self.grids):
                if i > 0:
                    g.band_connect_local(0, self.grids[i - 1])
                if i + 1 < len(self.grids):
                    g.band_connect_local(1, self.grids[i + 1])
            if len(self.grids) <= 16:
                for i, g in enumerate(self.grids):
                    g.band_rcx_connect_local(i, self.grids)

    def set_max_truncation(self, v):
        for g in self.grids:
            g.set_max_truncation(v)

    @property
    def bounds(self):
        return self.grids[0].bounds

    def free_footprint(self, *a):
        self.dirty = True  # (all columns: dirty_cols stays empty)
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

    def push(self, scan: Scan, sync: bool = True):
        box = self.grids[0].scan_box(scan)
        self.pushed = [band_reached(box, b, e) for (b, e) in self.rows]
        for g, mine in zip(self.grids, self.pushed):
            if mine:
                g.push_async(scan)
        self.dirty = True
        self.dirty_cols.add(box[0], box[2])
        if sync:
            self.sync_halos()

    def push_batch(self, scans, sync: bool = True):
        """tsdg_push_batch on every band any of the scans can reach."""
        boxes = [self.grids[0].scan_box(sc) for sc in scans]
        self.pushed = [any(band_reached(bx, b, e) for bx in boxes) for (b, e) in self.rows]
        for g, mine in zip(self.grids, self.pushed):
            if mine:
                g.push_batch_async(scans)
        self.dirty = True
        for bx in boxes:
            self.dirty_cols.add(bx[0], bx[2])
        if sync:
            self.sync_halos()

    def sync_halos(self):
        if not getattr(self, "dirty", False):
            return
        if self.peer:
            # every band's kernel waits for its neighbours': all of them are enqueued (on their own streams) before
            # anything synchronises
            cols = (self.dirty_cols.lo, self.dirty_cols.hi) if self.dirty_cols else (0, self.parts_x - 1)
            for g in self.grids:
                g.band_halo_sync(cols, cols)
            for g in self.grids:
                g.sync()
        else:
            self._exchange()
            for g in self.grids:
                g.band_push_finish()
        self.dirty = False
        self.dirty_cols.clear()

    def sync_flags(self):
        """Element-wise MAX of the bands' allocation flags (what DistBand does with an all-reduce)."""
        views = []
        for g in self.grids:
            g.sync()
            ptr, n = g.band_flags()
            views.append(device_tensor(ptr, n, "|u1", self.device))
        merged = torch.stack(views).max(dim=0).values
        for v in views:
            v.copy_(merged)
        torch.cuda.synchronize(self.device)

    def last_push_stats(self):
        sts = [g.last_push_stats() for g, mine in zip(self.grids, self.pushed) if mine]
        out = dict(sts[0])
        for k in ("cell_updates", "cell_visits", "fallback_cells", "active_tiles", "emptied_tiles", "newly_initialized"):
            out[k] = sum(s[k] for s in sts)
        return out

    def partition_states(self):
        """State and init weight of every partition, each taken from the band that owns it."""
        out = None
        for (b, e), g in zip(self.rows, self.grids):
            st, iw = g.partition_states()
            if out is None:
                out = (st.copy(), iw.copy())
            out[0][b * self.parts_x:e * self.parts_x] = st[b * self.parts_x:e * self.parts_x]
            out[1][b * self.parts_x:e * self.parts_x] = iw[b * self.parts_x:e * self.parts_x]
        return out

    def download_partition(self, p: int):
        py = p // self.parts_x
        for (b, e), g in zip(self.rows, self.grids):
            if b <= py < e:
                return g.download_partition(p)
        return None

    def raycast_mask(self, scan: Scan, rays_world, coords=None, normals=None):
        self.sync_halos()
        self.sync_flags()
        n = scan.n
        if self.peer and len(self.grids) <= 16 and n <= 2048:
            # the library's own merge over peer memory (tsdg_raycast_sharded_launch / _collect): every band's marching
            # kernel stores its events into every band's exchange block, a second kernel per band keeps the earliest
            for g in self.grids:
                g.raycast_sharded_launch(scan, rays_world)
            res = [g.raycast_sharded_collect(n) for g in self.grids]
            c0, n0, m0, k0 = res[0]
            for c1, n1, m1, k1 in res[1:]:  # every band holds the same, full result
                assert k1 == k0 and np.array_equal(m1, m0) and np.array_equal(c1, c0) and np.array_equal(n1, n0)
            if coords is not None:
                coords[m0 > 0] = c0[m0 > 0]
                c0 = coords
            if normals is not None:
                normals[m0 > 0] = n0[m0 > 0]
                n0 = normals
            return c0, n0, m0, k0
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

    def __init__(self, cell_size: float, layout_grid: int, device: int, transport: str = "peer"):
        """transport "peer": halo rows are stored straight into the neighbours' memory by one kernel
        (tsdg_band_halo_sync over CUDA IPC mappings, NVLink P2P); "nccl": batched isend/irecv + tsdg_band_push_finish."""
        import torch.distributed as dist
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.device = device
        parts_y = (1 << layout_grid) // 32
        self.rows = split_rows(parts_y, self.world)
        self.grid = capi.Grid(cell_size, 5, layout_grid, device=device, band=self.rows[self.rank])
        self.parts_x = parts_y
        self.dirty_lo, self.dirty_hi = DirtyColumns(), DirtyColumns()
        self.flags_dirty = False
        self._flags = None
        self._views = None
        self.transport = transport
        self.rcx = False
        self.dirty_rows = None  # partition rows whose allocation flags may have changed since the last merge
        # gloo (the CPU backend; also what the one-GPU multi-process parity test uses) reduces host tensors only
        self._host_collectives = dist.get_backend() != "nccl"
        if self._host_collectives and transport != "peer":
            raise ValueError("the NCCL halo transport needs the nccl backend; use transport='peer'")
        if transport == "peer" and self.world > 1:
            blobs = [None] * self.world
            dist.all_gather_object(blobs, self.grid.band_export())
            if self.rank > 0:
                self.grid.band_connect(0, blobs[self.rank - 1])
            if self.rank + 1 < self.world:
                self.grid.band_connect(1, blobs[self.rank + 1])
            # the ray-cast exchange connects every band with every band
            self.rcx = self.world <= 16
            if self.rcx:
                rblobs = [None] * self.world
                dist.all_gather_object(rblobs, self.grid.band_rcx_export())
                self.grid.band_rcx_connect(self.rank, rblobs)
            dist.barrier()

    def _all_reduce(self, t: torch.Tensor, op):
        """All-reduce of a device tensor on the current stream (NCCL), or through the host (gloo)."""
        if self._host_collectives:
            h = t.cpu()
            self.dist.all_reduce(h, op=op)
            t.copy_(h)
        else:
            self.dist.all_reduce(t, op=op)

    def _row_views(self):
        if self._views is None:
            v = {}
            for which in range(4):
                t, w, n = self.grid.band_row(which)
                v[which] = None if n == 0 else (device_tensor(t, n, "<f8", self.device), device_tensor(w, n, "<f8", self.device))
            self._views = v
        return self._views

    def note_scan(self, box) -> bool:
        """Book-keeping for one scan that EVERY rank calls with the same box: which boundaries it dirties.
        Returns whether this rank has to push it."""
        b, e = self.rows[self.rank]
        self.dirty_rows = (box[1], box[3]) if self.dirty_rows is None else (min(self.dirty_rows[0], box[1]), max(self.dirty_rows[1], box[3]))
        lower, upper = touched_boundaries(box, b, e, self.rank, self.world)
        if lower:
            self.dirty_lo.add(box[0], box[2])
        if upper:
            self.dirty_hi.add(box[0], box[2])
        return band_reached(box, b, e)

    def exchange(self, full: bool = False):
        """Boundary rows to the neighbouring bands, stream-ordered (no host synchronisation): the collective
        stream waits for the library's stream, the library's stream waits for the collective.  Only the dirty
        partition columns of each boundary travel (all of them with full=True)."""
        dist = self.dist
        v = self._row_views()
        cur = torch.cuda.current_stream(self.device)
        self.grid.stream_order(cur.cuda_stream, 0)
        ops = []
        stride = capi.TILE_STRIDE

        def cols(d):
            return (0, self.parts_x - 1) if full else (d.lo, d.hi)

        if self.rank > 0 and (full or self.dirty_lo):  # lowest row down, lower halo from below
            lo, hi = cols(self.dirty_lo)
            sl = slice(lo * stride, (hi + 1) * stride)
            for k in (0, 1):
                ops.append(dist.P2POp(dist.isend, v[0][k][sl], self.rank - 1))
                ops.append(dist.P2POp(dist.irecv, v[2][k][sl], self.rank - 1))
        if self.rank + 1 < self.world and (full or self.dirty_hi):
            lo, hi = cols(self.dirty_hi)
            sl = slice(lo * stride, (hi + 1) * stride)
            for k in (0, 1):
                ops.append(dist.P2POp(dist.isend, v[1][k][sl], self.rank + 1))
                ops.append(dist.P2POp(dist.irecv, v[3][k][sl], self.rank + 1))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()  # NCCL: orders the current stream after the transfer, does not block the host
        self.grid.stream_order(cur.cuda_stream, 1)
        return len(ops)

    def sync_halos(self, full: bool = False):
        """Bring the halos up to date (see the module docstring).  Collective: every rank calls it."""
        if not (full or self.dirty_lo or self.dirty_hi):
            return 0
        if self.transport == "peer":
            whole = (0, self.parts_x - 1)
            lo = whole if full else ((self.dirty_lo.lo, self.dirty_lo.hi) if self.dirty_lo else None)
            hi = whole if full else ((self.dirty_hi.lo, self.dirty_hi.hi) if self.dirty_hi else None)
            self.grid.band_halo_sync(lo, hi)
            n = 1
        else:
            n = self.exchange(full)
            self.grid.band_push_finish()
        self.dirty_lo.clear()
        self.dirty_hi.clear()
        return n

    def sync_flags(self):
        """Merge the bands' allocation flags (all-reduce MAX over the byte array) if a push happened since the last
        merge.  Collective: every rank calls it (ray casting does)."""
        if not self.flags_dirty:
            return
        if self._flags is None:
            ptr, n = self.grid.band_flags()
            self._flags = device_tensor(ptr, n, "|u1", self.device)
        # only the partition rows a scan could have reached since the last merge (every rank saw the same scans, hence
        # the same range); everything, when the range is unknown (fills, footprints)
        view = self._flags
        if self.dirty_rows is not None:
            r0, r1 = max(self.dirty_rows[0] - 1, 0), min(self.dirty_rows[1] + 1, self.parts_x - 1)
            view = self._flags[r0 * self.parts_x:(r1 + 1) * self.parts_x]
        cur = torch.cuda.current_stream(self.device)
        self.grid.stream_order(cur.cuda_stream, 0)
        self._all_reduce(view, self.dist.ReduceOp.MAX)
        self.grid.stream_order(cur.cuda_stream, 1)
        self.flags_dirty = False
        self.dirty_rows = None

    def _box(self, scan: Scan):
        """tsdg_scan_box, cached on the scan object (keyed by what the box depends on): a step of N robots asks for
        2N boxes on every rank."""
        key = (float(scan.pose[0, 2]), float(scan.pose[1, 2]), float(scan.spec.max_range))
        c = getattr(scan, "_box_cache", None)
        if c is None or c[0] != key:
            c = (key, self.grid.scan_box(scan))
            scan._box_cache = c
        return c[1]

    def push(self, scan: Scan, sync: bool = False):
        """Every rank calls this with the same scan; ranks the scan cannot reach skip it."""
        self.flags_dirty = True
        if self.note_scan(self._box(scan)):
            self.grid.push_async(scan)
        if sync:
            self.sync_halos()

    def push_batch(self, scans, sync: bool = False):
        """Every rank calls this with the same scans (tsdg_push_batch: the result of pushing them one by one)."""
        self.flags_dirty = True
        mine = [self.note_scan(self._box(sc)) for sc in scans]
        if any(mine):
            self.grid.push_batch_async(scans)
        if sync:
            self.sync_halos()

    def stage_and_note(self, scan: Scan) -> bool:
        """Staged variant: H2D of the scan if this rank has to push it; follow with push_staged()."""
        self.flags_dirty = True
        mine = self.note_scan(self._box(scan))
        if mine:
            self.grid.stage_scan(scan)
        return mine

    def push_staged(self):
        self.grid.push_staged()

    def raycast_mask(self, scan: Scan, rays_world):
        dist = self.dist
        n = scan.n
        self.sync_halos()
        self.sync_flags()
        if self.rcx and n <= 2048:
            # min-merge of the bands' crossings inside the library, over peer memory (no collective call here)
            return self.grid.raycast_mask_sharded(scan, rays_world)
        kp, pp = self.grid.raycast_band_keys(scan, rays_world)
        keys = device_tensor(kp, n, "<i8", self.device)
        payload = device_tensor(pp, 4 * n, "<f8", self.device).view(n, 4)
        mask, out, _ = merge_first_events(keys, payload, lambda t: self._all_reduce(t, dist.ReduceOp.MIN),
                                          lambda t: self._all_reduce(t, dist.ReduceOp.SUM))
        mask = mask.cpu().numpy().astype(np.uint8)
        out = out.cpu().numpy()
        return out[:, :2], out[:, 2:], mask, int(mask.sum())
