"""ctypes binding of libtsdslam_b200.so (include/tsdslam_b200.h): the reference-facing call path.

This module is plumbing only.  It fails loudly when the CUDA library is missing or when there is no
CUDA device: there is no CPU path behind it."""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import _build
from .scan import Hypothesis, PushStats, Scan, ScanStruct

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint32)
_bp = C.POINTER(C.c_ubyte)
_sp = C.POINTER(ScanStruct)
_hp = C.POINTER(Hypothesis)
_vpp = C.POINTER(C.c_void_p)

BAND_EXPORT_BYTES = 256  # include/tsdslam_b200.h TSD_BAND_EXPORT_BYTES
TILE_STRIDE = 1104  # doubles per partition and array (csrc/common.cuh TSD_TILE_STRIDE)

EXPORTS = [
    "tsd_last_error", "tsd_device_count", "tsd_kernel_launches", "tsd_invert3x3",
    "tsdg_create", "tsdg_create_band", "tsdg_band_push_finish", "tsdg_band_export", "tsdg_band_connect", "tsdg_band_connect_local", "tsdg_band_halo_sync",
    "tsdg_band_flags", "tsdg_scan_box", "tsdg_band_row", "tsdg_destroy", "tsdg_set_max_truncation", "tsdg_get_geometry",
    "tsdg_free_footprint", "tsdg_push", "tsdg_push_async", "tsdg_sync", "tsdg_stage_scan", "tsdg_push_staged", "tsdg_push_batch", "tsdg_push_batch_async", "tsdg_stage_batch",
    "tsdg_stream", "tsdg_stream_order", "tsdg_set_timing", "tsdg_set_update_filter", "tsdg_last_push_kernel_ms", "tsdg_last_push_stats", "tsdg_interpolate_bilinear", "tsdg_interpolate_normal",
    "tsdg_num_partitions", "tsdg_partition_states", "tsdg_download_partition", "tsdg_upload_partition", "tsdg_fill",
    "tsdg_raycast_mask", "tsdg_raycast", "tsdg_raycast_band_keys", "tsdg_last_raycast_steps",
    "tsdg_band_rcx_export", "tsdg_band_rcx_connect", "tsdg_band_rcx_connect_local", "tsdg_raycast_mask_sharded",
    "tsdg_raycast_sharded_launch", "tsdg_raycast_sharded_collect",
    "tsdg_axis_aligned_map", "tsdg_color_image", "tsdg_store", "tsdg_load", "tsds_prepare_scan",
    "tsdg_create_sharded", "tsdg_sharded_destroy", "tsdg_sharded_num_bands", "tsdg_sharded_band", "tsdg_sharded_set_max_truncation",
    "tsdg_sharded_free_footprint", "tsdg_sharded_push", "tsdg_sharded_push_batch", "tsdg_sharded_sync", "tsdg_sharded_last_push_stats",
    "tsdg_sharded_raycast_mask", "tsdg_sharded_interpolate_bilinear", "tsdg_sharded_partition_states", "tsdg_sharded_download_partition",
    "tsdg_localize", "icp_create", "icp_destroy", "icp_set_termination", "icp_set_max_iterations", "icp_run", "icp_pairs", "icp_set_trace", "icp_get_trace",
    "match_create", "match_destroy", "match_prepare", "match_rng", "match_score_tsd", "match_score_rnm", "match_score_pdf",
]


class TsdError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (building if the sources are newer) the CUDA library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if _build.needs_build():
        path = _build.build()
    L = C.CDLL(path)
    L.tsd_last_error.restype = C.c_char_p
    L.tsd_kernel_launches.restype = C.c_uint64
    L.tsd_invert3x3.argtypes = [_dp, _dp]
    L.tsdg_create.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, _vpp]
    L.tsdg_create_band.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vpp]
    L.tsdg_destroy.argtypes = [C.c_void_p]
    L.tsdg_band_push_finish.argtypes = [C.c_void_p]
    L.tsdg_scan_box.argtypes = [C.c_void_p, _sp, C.POINTER(C.c_int32)]
    L.tsdg_axis_aligned_map.argtypes = [C.c_void_p, _dp, C.c_uint32, _dp, C.POINTER(C.c_uint32), C.c_void_p]
    L.tsdg_color_image.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.tsdg_store.argtypes = [C.c_void_p, C.c_char_p]
    L.tsdg_load.argtypes = [C.c_char_p, C.c_int, _vpp]
    L.tsdg_band_flags.argtypes = [C.c_void_p, _vpp, C.POINTER(C.c_uint64)]
    L.tsdg_band_export.argtypes = [C.c_void_p, C.c_void_p]
    L.tsdg_band_connect.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.tsdg_band_connect_local.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.tsdg_band_halo_sync.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.tsdg_band_row.argtypes = [C.c_void_p, C.c_int, _vpp, _vpp, C.POINTER(C.c_uint64)]
    L.tsdg_raycast_band_keys.argtypes = [C.c_void_p, _sp, _dp, _vpp, _vpp]
    L.tsdg_band_rcx_export.argtypes = [C.c_void_p, C.c_void_p]
    L.tsdg_band_rcx_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.tsdg_band_rcx_connect_local.argtypes = [C.c_void_p, C.c_int, C.c_int, _vpp]
    L.tsdg_raycast_mask_sharded.argtypes = [C.c_void_p, _sp, _dp, _dp, _dp, _bp, _up]
    L.tsdg_raycast_sharded_launch.argtypes = [C.c_void_p, _sp, _dp]
    L.tsdg_raycast_sharded_collect.argtypes = [C.c_void_p, C.c_int32, _dp, _dp, _bp, _up]
    L.tsdg_set_max_truncation.argtypes = [C.c_void_p, C.c_double]
    L.tsdg_get_geometry.argtypes = [C.c_void_p, _ip, _ip, _ip] + [_dp] * 6
    L.tsdg_free_footprint.argtypes = [C.c_void_p] + [C.c_double] * 4
    L.tsdg_push.argtypes = [C.c_void_p, _sp]
    L.tsdg_push_async.argtypes = [C.c_void_p, _sp]
    L.tsdg_stage_scan.argtypes = [C.c_void_p, _sp]
    L.tsdg_push_batch.argtypes = [C.c_void_p, _sp, C.c_int32]
    L.tsdg_push_batch_async.argtypes = [C.c_void_p, _sp, C.c_int32]
    L.tsdg_stage_batch.argtypes = [C.c_void_p, _sp, C.c_int32]
    L.tsdg_push_staged.argtypes = [C.c_void_p]
    L.tsdg_sync.argtypes = [C.c_void_p]
    L.tsdg_stream.restype = C.c_void_p
    L.tsdg_stream.argtypes = [C.c_void_p]
    L.tsdg_stream_order.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.tsdg_set_timing.argtypes = [C.c_void_p, C.c_int]
    L.tsdg_set_update_filter.argtypes = [C.c_void_p, C.c_uint]
    L.tsds_prepare_scan.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.c_float, C.c_double, C.c_double, _dp, _dp, _bp, _dp,
                                    _bp, _dp, _up]
    L.tsdg_last_push_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.tsdg_last_push_stats.argtypes = [C.c_void_p, C.POINTER(PushStats)]
    L.tsdg_interpolate_bilinear.argtypes = [C.c_void_p, C.c_int32, _dp, _dp, _ip]
    L.tsdg_interpolate_normal.argtypes = [C.c_void_p, C.c_int32, _dp, _dp, _ip]
    L.tsdg_num_partitions.argtypes = [C.c_void_p, _ip]
    L.tsdg_partition_states.argtypes = [C.c_void_p, _ip, _dp]
    L.tsdg_download_partition.argtypes = [C.c_void_p, C.c_int32, _dp, _dp]
    L.tsdg_upload_partition.argtypes = [C.c_void_p, C.c_int32, _dp, _dp]
    L.tsdg_fill.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
    L.tsdg_raycast_mask.argtypes = [C.c_void_p, _sp, _dp, _dp, _dp, _bp, _up]
    L.tsdg_raycast.argtypes = [C.c_void_p, _sp, _dp, _dp, _dp, _up]
    L.tsdg_last_raycast_steps.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.tsdg_create_sharded.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), _vpp]
    L.tsdg_sharded_destroy.argtypes = [C.c_void_p]
    L.tsdg_sharded_num_bands.argtypes = [C.c_void_p]
    L.tsdg_sharded_band.argtypes = [C.c_void_p, C.c_int]
    L.tsdg_sharded_band.restype = C.c_void_p
    L.tsdg_sharded_set_max_truncation.argtypes = [C.c_void_p, C.c_double]
    L.tsdg_sharded_free_footprint.argtypes = [C.c_void_p] + [C.c_double] * 4
    L.tsdg_sharded_push.argtypes = [C.c_void_p, _sp]
    L.tsdg_sharded_push_batch.argtypes = [C.c_void_p, _sp, C.c_int32]
    L.tsdg_sharded_sync.argtypes = [C.c_void_p]
    L.tsdg_sharded_last_push_stats.argtypes = [C.c_void_p, C.POINTER(PushStats)]
    L.tsdg_sharded_raycast_mask.argtypes = [C.c_void_p, _sp, _dp, _dp, _dp, _bp, _up]
    L.tsdg_sharded_interpolate_bilinear.argtypes = [C.c_void_p, C.c_int32, _dp, _dp, _ip]
    L.tsdg_sharded_partition_states.argtypes = [C.c_void_p, _ip, _dp]
    L.tsdg_sharded_download_partition.argtypes = [C.c_void_p, C.c_int32, _dp, _dp]
    L.icp_create.argtypes = [C.c_uint32, C.c_double, C.c_double, C.c_uint32, _dp, C.c_int, _vpp]
    L.icp_destroy.argtypes = [C.c_void_p]
    L.icp_run.argtypes = [C.c_void_p, _dp, _dp, C.c_int32, _dp, C.c_int32, _dp, _dp, _dp, _dp, _up, _up, _ip]
    L.tsdg_localize.argtypes = [C.c_void_p, C.c_void_p, _sp, _dp, _dp, C.c_int32, _dp, _dp, _dp, _up, _up, _ip, _up]
    L.icp_pairs.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, C.c_int32, _dp, _up, _up, _dp, _up]
    L.icp_set_termination.argtypes = [C.c_void_p, C.c_double, C.c_uint32]
    L.icp_set_max_iterations.argtypes = [C.c_void_p, C.c_uint32]
    L.icp_set_trace.argtypes = [C.c_void_p, C.c_int]
    L.icp_get_trace.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _up, _up, _ip, _dp, _dp, _ip]
    L.match_create.argtypes = [C.c_int, _vpp]
    L.match_destroy.argtypes = [C.c_void_p]
    L.match_prepare.argtypes = [C.c_void_p, C.c_int32, _dp, _bp, _dp, _bp, C.c_int32, C.c_uint32, C.c_uint32, C.c_double, C.c_double,
                                C.c_uint64, C.POINTER(MatchPrepStruct)]
    L.match_rng.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    L.match_rng.restype = C.c_uint64
    L.match_score_tsd.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, _hp, C.c_int32, _dp, _dp, _dp, _dp, C.c_double,
                                  C.c_int32, _dp, _dp, C.c_double, _dp, _ip, _dp]
    L.match_score_rnm.argtypes = [C.c_void_p, C.c_int32, _hp, C.c_int32, _dp, _dp, _dp, _dp, C.c_double, C.c_int32, _dp,
                                  _dp, C.c_int32, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint32, _ip,
                                  _ip, _dp, _ip, _dp]
    L.match_score_pdf.argtypes = [C.c_void_p, C.c_int32, _hp, C.c_int32, _dp, _dp, _dp, _dp, C.c_double, C.c_int32, _dp,
                                  C.c_int32, _dp, _dp, _dp, _dp, _ip, _ip, _dp]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise TsdError(f"libtsdslam_b200 error {rc}: {lib().tsd_last_error().decode()}")


def device_count() -> int:
    return int(lib().tsd_device_count())


def kernel_launches() -> int:
    return int(lib().tsd_kernel_launches())


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def invert3x3(m) -> np.ndarray:
    m = _f64(m)
    out = np.empty((3, 3))
    check(lib().tsd_invert3x3(_d(m), _d(out)))
    return out


class ScanBatch:
    """Scans laid out as the contiguous tsd_scan_t array tsdg_push_batch takes (built once, reused every cycle)."""

    def __init__(self, scans):
        self.scans = list(scans)  # keeps the numpy buffers alive
        self.array = (ScanStruct * len(self.scans))()
        for i, sc in enumerate(self.scans):
            self.array[i] = sc.struct

    def __len__(self):
        return len(self.scans)

    def __iter__(self):
        return iter(self.scans)


class Grid:
    """obvious::TsdGrid on the device."""

    def __init__(self, cell_size: float, layout_partition: int, layout_grid: int, device: int = 0, band=None, handle=None):
        h = C.c_void_p()
        if handle is not None:
            h = handle
        elif band is None:
            check(lib().tsdg_create(cell_size, layout_partition, layout_grid, device, C.byref(h)))
        else:
            check(lib().tsdg_create_band(cell_size, layout_partition, layout_grid, device, band[0], band[1], C.byref(h)))
        self.h = h
        self.cell_size = cell_size
        self.dim = 1 << layout_partition
        self.cells = 1 << layout_grid
        self.parts_per_side = self.cells // self.dim
        self.n_partitions = self.parts_per_side ** 2

    @classmethod
    def load(cls, path: str, device: int = 0):
        """TsdGrid(path, FILE_SOURCE): a grid from the reference's checkpoint format (tsdg_load)."""
        h = C.c_void_p()
        check(lib().tsdg_load(path.encode(), device, C.byref(h)))
        with open(path) as f:
            cs, lp, lg = float(f.readline()), int(f.readline()), int(f.readline())
        return cls(cs, lp, lg, device=device, handle=h)

    def store(self, path: str) -> bool:
        """TsdGrid::storeGrid"""
        check(lib().tsdg_store(self.h, path.encode()))
        return True

    def close(self):
        if getattr(self, "h", None):
            if not getattr(self, "_borrowed", False):
                lib().tsdg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_max_truncation(self, v):
        check(lib().tsdg_set_max_truncation(self.h, v))

    @property
    def bounds(self):
        v = [C.c_double() for _ in range(4)]
        check(lib().tsdg_get_geometry(self.h, None, None, None, None, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]),
                                      C.byref(v[3]), None))
        return tuple(x.value for x in v)

    def free_footprint(self, cx, cy, w, h) -> bool:
        rc = lib().tsdg_free_footprint(self.h, cx, cy, w, h)
        if rc == -5:
            return False
        check(rc)
        return True

    def push(self, scan: Scan):
        check(lib().tsdg_push(self.h, scan.byref()))

    def push_async(self, scan: Scan):
        check(lib().tsdg_push_async(self.h, scan.byref()))

    @staticmethod
    def _scan_array(scans):
        if isinstance(scans, ScanBatch):
            return scans.array
        arr = (ScanStruct * len(scans))()
        for i, sc in enumerate(scans):
            arr[i] = sc.struct
        return arr

    def push_batch(self, scans):
        """The scans in order, as len(scans) pushes would; pairs of scans share their launches (tsdg_push_batch).
        `scans`: a list of Scan or a prepared ScanBatch."""
        check(lib().tsdg_push_batch(self.h, self._scan_array(scans), len(scans)))

    def push_batch_async(self, scans):
        check(lib().tsdg_push_batch_async(self.h, self._scan_array(scans), len(scans)))

    def stage_batch(self, scans):
        check(lib().tsdg_stage_batch(self.h, self._scan_array(scans), len(scans)))

    def stage_scan(self, scan: Scan):
        check(lib().tsdg_stage_scan(self.h, scan.byref()))

    def push_staged(self):
        check(lib().tsdg_push_staged(self.h))

    def sync(self):
        check(lib().tsdg_sync(self.h))

    @property
    def stream_ptr(self) -> int:
        return int(lib().tsdg_stream(self.h) or 0)

    def stream_order(self, other_stream_ptr: int, direction: int):
        check(lib().tsdg_stream_order(self.h, C.c_void_p(other_stream_ptr), direction))

    def last_push_stats(self):
        st = PushStats()
        check(lib().tsdg_last_push_stats(self.h, C.byref(st)))
        return st.as_dict()

    def set_timing(self, enable: bool = True):
        check(lib().tsdg_set_timing(self.h, 1 if enable else 0))

    def prepare_scan(self, ranges_f32, spec, rays_local, scale: float = 1.0, scene=None):
        """Scan pre-processing on the device (tsds_prepare_scan): returns data, mask, scene (n x 2), scene mask and
        the compacted valid scene points."""
        r = np.ascontiguousarray(ranges_f32, dtype=np.float32)
        n = len(r)
        rays = _f64(rays_local)
        data = np.empty(n)
        mask = np.empty(n, dtype=np.uint8)
        scene = np.zeros((n, 2)) if scene is None else scene
        smask = np.empty(n, dtype=np.uint8)
        valid = np.empty((n, 2))
        cnt = C.c_uint32()
        check(lib().tsds_prepare_scan(self.h, n, r.ctypes.data_as(C.POINTER(C.c_float)), scale, spec.max_range, spec.angular_res,
                                      _d(rays), _d(data), mask.ctypes.data_as(_bp), _d(scene), smask.ctypes.data_as(_bp),
                                      _d(valid), C.byref(cnt)))
        return data, mask, scene, smask, valid[:cnt.value].copy()

    def set_update_filter(self, mask: int):
        """Measurement aid: bit 0 = the update kernel skips K2 (addTsd) work, bit 1 = K3 (increaseEmptiness) work."""
        check(lib().tsdg_set_update_filter(self.h, mask))

    def last_push_kernel_ms(self):
        ms = (C.c_float * 4)()
        check(lib().tsdg_last_push_kernel_ms(self.h, ms))
        return dict(classify=ms[0], update=ms[1], borders=ms[2], total=ms[3])

    def partition_states(self):
        st = np.empty(self.n_partitions, dtype=np.int32)
        iw = np.empty(self.n_partitions)
        check(lib().tsdg_partition_states(self.h, st.ctypes.data_as(_ip), _d(iw)))
        return st, iw

    def download_partition(self, p: int):
        n = (self.dim + 1) ** 2
        tsd = np.empty(n)
        w = np.empty(n)
        if lib().tsdg_download_partition(self.h, p, _d(tsd), _d(w)) != 0:
            return None
        return tsd.reshape(self.dim + 1, -1), w.reshape(self.dim + 1, -1)

    def upload_partition(self, p: int, tsd, w):
        tsd, w = _f64(tsd), _f64(w)
        check(lib().tsdg_upload_partition(self.h, p, _d(tsd), _d(w)))

    def fill(self, tsd, weight, only_uninitialized: bool = False):
        check(lib().tsdg_fill(self.h, tsd, weight, 1 if only_uninitialized else 0))

    def interpolate_bilinear(self, xy):
        xy = _f64(xy)
        tsd = np.empty(len(xy))
        st = np.empty(len(xy), dtype=np.int32)
        check(lib().tsdg_interpolate_bilinear(self.h, len(xy), _d(xy), _d(tsd), st.ctypes.data_as(_ip)))
        return tsd, st

    def interpolate_normal(self, xy):
        xy = _f64(xy)
        nn = np.empty((len(xy), 2))
        ok = np.empty(len(xy), dtype=np.int32)
        check(lib().tsdg_interpolate_normal(self.h, len(xy), _d(xy), _d(nn), ok.ctypes.data_as(_ip)))
        return nn, ok

    def raycast_mask(self, scan: Scan, rays_world, coords=None, normals=None):
        n = scan.n
        rays = _f64(rays_world)
        coords = np.zeros((n, 2)) if coords is None else coords
        normals = np.zeros((n, 2)) if normals is None else normals
        mask = np.zeros(n, dtype=np.uint8)
        cnt = C.c_uint32()
        check(lib().tsdg_raycast_mask(self.h, scan.byref(), _d(rays), _d(coords), _d(normals), mask.ctypes.data_as(_bp),
                                      C.byref(cnt)))
        return coords, normals, mask, int(cnt.value)

    def raycast(self, scan: Scan, rays_world):
        n = scan.n
        rays = _f64(rays_world)
        coords = np.zeros(2 * n)
        normals = np.zeros(2 * n)
        cnt = C.c_uint32()
        check(lib().tsdg_raycast(self.h, scan.byref(), _d(rays), _d(coords), _d(normals), C.byref(cnt)))
        k = int(cnt.value)
        return coords[:k].reshape(-1, 2), normals[:k].reshape(-1, 2)

    def axis_map(self, with_normals: bool = False, occupied=None, cap_points=None):
        """RayCastAxisAligned2D::calcCoords: (coords (k, 2), normals (k, 2) or None, occupied int8[cells*cells])."""
        cells = self.cells
        cap = cap_points or (cells * cells) // 2
        coords = np.zeros(2 * cap)
        normals = np.full(2 * cap, np.nan) if with_normals else None
        occ = np.full(cells * cells, -1, dtype=np.int8) if occupied is None else occupied
        cnt = C.c_uint32()
        check(lib().tsdg_axis_aligned_map(self.h, _d(coords), cap, _d(normals) if with_normals else None, C.byref(cnt),
                                          occ.ctypes.data_as(C.c_void_p)))
        k = int(cnt.value) // 2
        return coords[:2 * k].reshape(-1, 2), (normals[:2 * k].reshape(-1, 2) if with_normals else None), occ

    def color_image(self, width: int, height: int):
        img = np.zeros(3 * width * height, dtype=np.uint8)
        check(lib().tsdg_color_image(self.h, img.ctypes.data_as(C.c_void_p), width, height))
        return img.reshape(height, width, 3)

    def band_push_finish(self):
        check(lib().tsdg_band_push_finish(self.h))

    def band_export(self) -> bytes:
        """Opaque description of this band (CUDA IPC handles) for the neighbouring processes."""
        blob = C.create_string_buffer(BAND_EXPORT_BYTES)
        check(lib().tsdg_band_export(self.h, blob))
        return blob.raw

    def band_connect(self, side: int, blob: bytes):
        """side 0: the band below, 1: the band above (a blob from that band's band_export)."""
        buf = C.create_string_buffer(blob, BAND_EXPORT_BYTES)
        check(lib().tsdg_band_connect(self.h, side, buf))

    def band_connect_local(self, side: int, other: "Grid"):
        check(lib().tsdg_band_connect_local(self.h, side, other.h))

    def band_halo_sync(self, lo=None, hi=None):
        """lo / hi: (px0, px1) dirty partition columns at the boundary below / above; None = untouched."""
        l = lo if lo is not None else (0, -1)
        h = hi if hi is not None else (0, -1)
        check(lib().tsdg_band_halo_sync(self.h, int(l[0]), int(l[1]), int(h[0]), int(h[1])))

    def band_flags(self):
        """(device pointer, count) of the allocation flags of all partitions."""
        f, n = C.c_void_p(), C.c_uint64()
        check(lib().tsdg_band_flags(self.h, C.byref(f), C.byref(n)))
        return int(f.value or 0), int(n.value)

    def scan_box(self, scan: Scan):
        """Inclusive partition box (px0, py0, px1, py1) the scan can touch (tsdg_scan_box)."""
        box = (C.c_int32 * 4)()
        check(lib().tsdg_scan_box(self.h, scan.byref(), box))
        return tuple(int(v) for v in box)

    def band_row(self, which: int):
        """(tsd_ptr, weight_ptr, count_doubles) of a boundary / halo partition row; (0, 0, 0) if absent."""
        t, w, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        check(lib().tsdg_band_row(self.h, which, C.byref(t), C.byref(w), C.byref(n)))
        return int(t.value or 0), int(w.value or 0), int(n.value)

    def band_rcx_export(self) -> bytes:
        buf = C.create_string_buffer(BAND_EXPORT_BYTES)
        check(lib().tsdg_band_rcx_export(self.h, buf))
        return bytes(buf.raw)

    def band_rcx_connect(self, rank: int, blobs):
        """blobs: the tsdg_band_rcx_export blobs of ALL bands, ordered by rank."""
        raw = b"".join(blobs)
        check(lib().tsdg_band_rcx_connect(self.h, rank, len(blobs), C.create_string_buffer(raw, len(raw))))

    def band_rcx_connect_local(self, rank: int, grids):
        arr = (C.c_void_p * len(grids))(*[g.h for g in grids])
        check(lib().tsdg_band_rcx_connect_local(self.h, rank, len(grids), arr))

    def raycast_mask_sharded(self, scan: Scan, rays_world, coords=None, normals=None):
        """Collective over the bands of a sharded grid (one call per band / process): the full ray-cast result."""
        n = scan.n
        rays = _f64(rays_world)
        coords = np.zeros((n, 2)) if coords is None else coords
        normals = np.zeros((n, 2)) if normals is None else normals
        mask = np.zeros(n, dtype=np.uint8)
        cnt = C.c_uint32()
        check(lib().tsdg_raycast_mask_sharded(self.h, scan.byref(), _d(rays), _d(coords), _d(normals), mask.ctypes.data_as(_bp),
                                               C.byref(cnt)))
        return coords, normals, mask, int(cnt.value)

    def raycast_sharded_launch(self, scan: Scan, rays_world):
        check(lib().tsdg_raycast_sharded_launch(self.h, scan.byref(), _d(_f64(rays_world))))

    def raycast_sharded_collect(self, n: int, coords=None, normals=None):
        coords = np.zeros((n, 2)) if coords is None else coords
        normals = np.zeros((n, 2)) if normals is None else normals
        mask = np.zeros(n, dtype=np.uint8)
        cnt = C.c_uint32()
        check(lib().tsdg_raycast_sharded_collect(self.h, n, _d(coords), _d(normals), mask.ctypes.data_as(_bp), C.byref(cnt)))
        return coords, normals, mask, int(cnt.value)

    def raycast_band_keys(self, scan: Scan, rays_world):
        """Device pointers (keys u64[n], payload f64[4n]) of this band's first events."""
        rays = _f64(rays_world)
        k, p = C.c_void_p(), C.c_void_p()
        check(lib().tsdg_raycast_band_keys(self.h, scan.byref(), _d(rays), C.byref(k), C.byref(p)))
        return int(k.value), int(p.value)

    def raycast_steps(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(lib().tsdg_last_raycast_steps(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


class ShardedGrid:
    """tsd_sharded_t: one TsdGrid in n_bands bands of partition rows, one handle, one process (csrc/sharded.cu)."""

    def __init__(self, cell_size: float, layout_partition: int, layout_grid: int, n_bands: int, devices=None):
        self.h = C.c_void_p()
        dev = None if devices is None else (C.c_int * n_bands)(*devices)
        check(lib().tsdg_create_sharded(cell_size, layout_partition, layout_grid, n_bands, dev, C.byref(self.h)))
        self.cell_size = cell_size
        self.n_bands = lib().tsdg_sharded_num_bands(self.h)
        self.parts_x = (1 << layout_grid) // 32
        self.n_parts = self.parts_x * self.parts_x

    def close(self):
        if getattr(self, "h", None):
            lib().tsdg_sharded_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def band(self, i: int) -> "Grid":
        """A view of band i's own handle (not owned: do not close)."""
        g = Grid.__new__(Grid)
        g.h = C.c_void_p(lib().tsdg_sharded_band(self.h, i))
        g._borrowed = True
        return g

    def set_max_truncation(self, v):
        check(lib().tsdg_sharded_set_max_truncation(self.h, v))

    def free_footprint(self, cx, cy, w, h) -> bool:
        rc = lib().tsdg_sharded_free_footprint(self.h, cx, cy, w, h)
        if rc == -5:
            return False
        check(rc)
        return True

    def push(self, scan: Scan):
        check(lib().tsdg_sharded_push(self.h, scan.byref()))

    def push_batch(self, scans):
        arr = Grid._scan_array(scans)
        check(lib().tsdg_sharded_push_batch(self.h, arr, len(arr)))

    def sync(self):
        check(lib().tsdg_sharded_sync(self.h))

    def last_push_stats(self):
        st = PushStats()
        check(lib().tsdg_sharded_last_push_stats(self.h, C.byref(st)))
        return st.as_dict()

    def raycast_mask(self, scan: Scan, rays_world, coords=None, normals=None):
        n = scan.n
        rays = _f64(rays_world)
        coords = np.zeros((n, 2)) if coords is None else coords
        normals = np.zeros((n, 2)) if normals is None else normals
        mask = np.zeros(n, dtype=np.uint8)
        cnt = C.c_uint32()
        check(lib().tsdg_sharded_raycast_mask(self.h, scan.byref(), _d(rays), _d(coords), _d(normals), mask.ctypes.data_as(_bp),
                                              C.byref(cnt)))
        return coords, normals, mask, int(cnt.value)

    def interpolate_bilinear(self, xy):
        xy = _f64(xy)
        tsd = np.empty(len(xy))
        st = np.empty(len(xy), dtype=np.int32)
        check(lib().tsdg_sharded_interpolate_bilinear(self.h, len(xy), _d(xy), _d(tsd), st.ctypes.data_as(_ip)))
        return tsd, st

    def partition_states(self):
        st = np.empty(self.n_parts, dtype=np.int32)
        iw = np.empty(self.n_parts)
        check(lib().tsdg_sharded_partition_states(self.h, st.ctypes.data_as(_ip), _d(iw)))
        return st, iw

    def download_partition(self, p: int):
        t = np.empty(33 * 33)
        w = np.empty(33 * 33)
        if lib().tsdg_sharded_download_partition(self.h, p, _d(t), _d(w)) != 0:
            return None
        return t.reshape(33, 33), w.reshape(33, 33)


class Icp:
    """obvious::Icp wired as ThreadLocalize.cpp:210-225, on the device."""

    def __init__(self, max_iterations, dist_max, dist_min, bounds, device: int = 0):
        b = _f64(bounds)
        self.max_iterations = max_iterations
        h = C.c_void_p()
        check(lib().icp_create(max_iterations, dist_max, dist_min, (max_iterations - 10) & 0xFFFFFFFF, _d(b), device,
                               C.byref(h)))
        self.h = h
        self._tracing = False

    def set_trace(self, enable: bool = True):
        check(lib().icp_set_trace(self.h, 1 if enable else 0))
        self._tracing = enable

    def close(self):
        if getattr(self, "h", None):
            lib().icp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, model, normals, scene, pose, Tinit44=None):
        model, normals, scene, pose = _f64(model), _f64(normals), _f64(scene), _f64(pose)
        Ti = None if Tinit44 is None else _f64(Tinit44)
        T = np.empty((3, 3))
        mse, pairs, its, st = C.c_double(), C.c_uint32(), C.c_uint32(), C.c_int32()
        check(lib().icp_run(self.h, _d(model), _d(normals), len(model), _d(scene), len(scene), _d(pose),
                            None if Ti is None else _d(Ti), _d(T), C.byref(mse), C.byref(pairs), C.byref(its), C.byref(st)))
        return T, mse.value, pairs.value, its.value, st.value

    def localize(self, grid, scan, rays_world, scene, Tinit44=None):
        """tsdg_localize: ray cast from scan.pose + maskMatrix + Icp::iterate with the model kept on the device.
        Returns (T, mse, pairs, iterations, state, n_model)."""
        rays, scene = _f64(rays_world), _f64(scene)
        Ti = None if Tinit44 is None else _f64(Tinit44)
        T = np.empty((3, 3))
        mse, pairs, its, st, nm = C.c_double(), C.c_uint32(), C.c_uint32(), C.c_int32(), C.c_uint32()
        check(lib().tsdg_localize(grid.h, self.h, scan.byref(), _d(rays), _d(scene), len(scene), None if Ti is None else _d(Ti),
                                  _d(T), C.byref(mse), C.byref(pairs), C.byref(its), C.byref(st), C.byref(nm)))
        return T, mse.value, pairs.value, its.value, st.value, nm.value

    def pairs(self, model, scene, pose):
        """PairAssignment::determinePairs on its own (icp_pairs): (model indices, scene indices, squared distances)."""
        model, scene, pose = _f64(model), _f64(scene), _f64(pose)
        cap = max(min(len(model), len(scene)), 1)
        pm = np.zeros(cap, dtype=np.uint32)
        ps = np.zeros(cap, dtype=np.uint32)
        d = np.zeros(cap)
        n = C.c_uint32()
        check(lib().icp_pairs(self.h, _d(model), len(model), _d(scene), len(scene), _d(pose), pm.ctypes.data_as(_up),
                              ps.ctypes.data_as(_up), _d(d), C.byref(n)))
        return pm[:n.value].copy(), ps[:n.value].copy(), d[:n.value].copy()

    def trace(self, cap):
        mi = self.max_iterations
        pm = np.zeros((mi, cap), dtype=np.uint32)
        ps = np.zeros((mi, cap), dtype=np.uint32)
        pc = np.zeros(mi, dtype=np.int32)
        mse = np.zeros(mi)
        Tf = np.zeros((mi, 4, 4))
        nit = C.c_int32()
        check(lib().icp_get_trace(self.h, mi, cap, pm.ctypes.data_as(_up), ps.ctypes.data_as(_up), pc.ctypes.data_as(_ip),
                                  _d(mse), _d(Tf), C.byref(nit)))
        return nit.value, pm, ps, pc, mse, Tf


def _hyps(h):
    h = np.ascontiguousarray(h, dtype=np.int32).reshape(-1, 2)
    return h, h.ctypes.data_as(_hp)


class MatchPrepStruct(C.Structure):
    """tsd_match_prep_t (include/tsdslam_b200.h)"""
    _fields_ = [("n", C.c_int32), ("n_hyp", C.c_int32), ("n_control", C.c_int32), ("n_trials", C.c_int32),
                ("n_valid_m", C.c_int32), ("n_valid_s", C.c_int32), ("span", C.c_int32),
                ("phi_max", C.c_double), ("theta_min", C.c_double), ("theta_max", C.c_double),
                ("hyps", C.c_void_p), ("model", C.c_void_p), ("scene", C.c_void_p), ("phi_m", C.c_void_p), ("phi_s", C.c_void_p),
                ("mask_m_pca", C.c_void_p), ("mask_s_pca", C.c_void_p), ("idx_m_valid", C.c_void_p), ("idx_s_valid", C.c_void_p),
                ("idx_control", C.c_void_p), ("idx_trials", C.c_void_p), ("control", C.c_void_p), ("phi_control", C.c_void_p),
                ("model_valid", C.c_void_p), ("phi_valid", C.c_void_p), ("model_angles", C.c_void_p), ("model_dists", C.c_void_p)]


class MatchPrep:
    """Result of match_prepare: numpy copies for inspection + the raw struct, whose pointers name the device-resident set."""

    def __init__(self, raw: MatchPrepStruct, copy: bool = True):
        self.raw = raw
        for f in ("n", "n_hyp", "n_control", "n_trials", "n_valid_m", "n_valid_s", "span", "phi_max", "theta_min", "theta_max"):
            setattr(self, f, getattr(raw, f))
        if not copy:
            return

        def arr(ptr, count, dtype):
            if count <= 0 or not ptr:
                return np.zeros(0, dtype=dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype, count=count).copy()
        n, c, vm = raw.n, raw.n_control, raw.n_valid_m
        self.hyps = arr(raw.hyps, 2 * raw.n_hyp, np.int32).reshape(-1, 2)
        self.phi_m, self.phi_s = arr(raw.phi_m, n, np.float64), arr(raw.phi_s, n, np.float64)
        self.mask_m_pca, self.mask_s_pca = arr(raw.mask_m_pca, n, np.uint8), arr(raw.mask_s_pca, n, np.uint8)
        self.idx_m_valid, self.idx_s_valid = arr(raw.idx_m_valid, vm, np.int32), arr(raw.idx_s_valid, raw.n_valid_s, np.int32)
        self.idx_control, self.idx_trials = arr(raw.idx_control, c, np.int32), arr(raw.idx_trials, raw.n_trials, np.int32)
        self.control = arr(raw.control, 3 * c, np.float64).reshape(3, c)
        self.phi_control = arr(raw.phi_control, c, np.float64)
        self.model_valid = arr(raw.model_valid, 2 * vm, np.float64).reshape(-1, 2)
        self.phi_valid = arr(raw.phi_valid, vm, np.float64)
        self.model_angles, self.model_dists = arr(raw.model_angles, vm, np.float64), arr(raw.model_dists, vm, np.float64)


def match_rng(seed: int, stream: int, index: int) -> int:
    return int(lib().match_rng(seed, stream, index))


class Matcher:
    """Hypothesis scorers of TSD_PDFMatching / RandomNormalMatching / PDFMatching on the device."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        check(lib().match_create(device, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().match_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def prepare(self, M, mask_m, S, mask_s, pca_search_range=10, size_control_set=360, trials=30, phi_max=math.pi / 4,
                resolution=0.0, seed=1, copy=True) -> MatchPrep:
        """match_prepare: the matchers' pre-processing on the device (counter-based random numbers)."""
        M, S = _f64(M), _f64(S)
        mm, ms = np.ascontiguousarray(mask_m, dtype=np.uint8), np.ascontiguousarray(mask_s, dtype=np.uint8)
        raw = MatchPrepStruct()
        check(lib().match_prepare(self.h, len(M), _d(M), mm.ctypes.data_as(_bp), _d(S), ms.ctypes.data_as(_bp), pca_search_range,
                                  size_control_set, trials, phi_max, resolution, seed, C.byref(raw)))
        return MatchPrep(raw, copy)

    def score_prepared(self, which: str, prep: MatchPrep, grid: "Grid" = None, t_sensor=None, zrand=0.05, scale_distance=1.0,
                       scale_orientation=1.0, params=None):
        """Scores the device-resident prepared set (no upload of it): which = "tsd" | "rnm" | "pdf".  Returns (per-hypothesis
        array(s)..., best, T)."""
        r = prep.raw
        nh = r.n_hyp
        best = C.c_int32()
        T = np.empty((3, 3))
        cast = lambda p, t: C.cast(C.c_void_p(p), t)
        common = (nh, cast(r.hyps, _hp), r.n, cast(r.model, _dp), cast(r.scene, _dp), cast(r.phi_m, _dp), cast(r.phi_s, _dp), r.phi_max,
                  r.n_control, cast(r.control, _dp))
        if which == "tsd":
            score = np.empty(max(nh, 1))
            ts = _f64(t_sensor)
            check(lib().match_score_tsd(self.h, grid.h, *common, _d(ts), zrand, _d(score), C.byref(best), _d(T)))
            return score[:nh], best.value, T
        if which == "rnm":
            cnt = np.empty(max(nh, 1), dtype=np.int32)
            mx = np.empty(max(nh, 1), dtype=np.int32)
            err = np.empty(max(nh, 1))
            check(lib().match_score_rnm(self.h, *common, cast(r.phi_control, _dp), r.n_valid_m, cast(r.model_valid, _dp),
                                        cast(r.phi_valid, _dp), r.theta_min, r.theta_max, scale_distance, scale_orientation,
                                        r.n_control // 3, cnt.ctypes.data_as(_ip), mx.ctypes.data_as(_ip), _d(err), C.byref(best), _d(T)))
            return cnt[:nh], mx[:nh], err[:nh], best.value, T
        if which == "pdf":
            prob = np.empty(max(nh, 1))
            fov = np.empty(max(nh, 1), dtype=np.int32)
            pr = _f64(params)
            check(lib().match_score_pdf(self.h, *common, r.n_valid_m, cast(r.model_angles, _dp), cast(r.model_dists, _dp), _d(pr),
                                        _d(prob), fov.ctypes.data_as(_ip), C.byref(best), _d(T)))
            return prob[:nh], fov[:nh], best.value, T
        raise ValueError(which)

    def score_tsd(self, grid: Grid, hyps, M, S, phi_m, phi_s, phi_max, control, t_sensor, zrand):
        h, hp = _hyps(hyps)
        M, S, phi_m, phi_s, control, t_sensor = map(_f64, (M, S, phi_m, phi_s, control, t_sensor))
        score = np.empty(len(h))
        best = C.c_int32()
        T = np.empty((3, 3))
        check(lib().match_score_tsd(self.h, grid.h, len(h), hp, len(M), _d(M), _d(S), _d(phi_m), _d(phi_s), phi_max,
                                    control.shape[1], _d(control), _d(t_sensor), zrand, _d(score), C.byref(best), _d(T)))
        return score, best.value, T

    def score_rnm(self, hyps, M, S, phi_m, phi_s, phi_max, control, phi_control, model_valid, phi_valid, theta_min,
                  theta_max, scale_distance, scale_orientation, cnt_thresh):
        h, hp = _hyps(hyps)
        M, S, phi_m, phi_s, control, phi_control, model_valid, phi_valid = map(
            _f64, (M, S, phi_m, phi_s, control, phi_control, model_valid, phi_valid))
        cnt = np.empty(len(h), dtype=np.int32)
        mx = np.empty(len(h), dtype=np.int32)
        err = np.empty(len(h))
        best = C.c_int32()
        T = np.empty((3, 3))
        check(lib().match_score_rnm(self.h, len(h), hp, len(M), _d(M), _d(S), _d(phi_m), _d(phi_s), phi_max,
                                    control.shape[1], _d(control), _d(phi_control), len(model_valid), _d(model_valid),
                                    _d(phi_valid), theta_min, theta_max, scale_distance, scale_orientation, cnt_thresh,
                                    cnt.ctypes.data_as(_ip), mx.ctypes.data_as(_ip), _d(err), C.byref(best), _d(T)))
        return cnt, mx, err, best.value, T

    def score_pdf(self, hyps, M, S, phi_m, phi_s, phi_max, control, model_angles, model_dists, params):
        h, hp = _hyps(hyps)
        M, S, phi_m, phi_s, control, model_angles, model_dists, params = map(
            _f64, (M, S, phi_m, phi_s, control, model_angles, model_dists, params))
        prob = np.empty(len(h))
        fov = np.empty(len(h), dtype=np.int32)
        best = C.c_int32()
        T = np.empty((3, 3))
        check(lib().match_score_pdf(self.h, len(h), hp, len(M), _d(M), _d(S), _d(phi_m), _d(phi_s), phi_max,
                                    control.shape[1], _d(control), len(model_angles), _d(model_angles), _d(model_dists),
                                    _d(params), _d(prob), fov.ctypes.data_as(_ip), C.byref(best), _d(T)))
        return prob, fov, best.value, T
