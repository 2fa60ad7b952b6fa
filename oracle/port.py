"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libohm_port.so (oracle/port/*.c).

Same Python-level interface as ohm_tsd_slam_b200.capi so that the parity tests run one harness against
the port (CPU) and the CUDA library.  The library is (re)built on demand with `make -C oracle port`
(plain gcc; works on the GPU box as well).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from ohm_tsd_slam_b200.scan import Hypothesis, PushStats, Scan, ScanStruct

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libohm_port.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint32)
_bp = C.POINTER(C.c_ubyte)
_sp = C.POINTER(ScanStruct)
_hp = C.POINTER(Hypothesis)


def build(force: bool = False):
    srcs = [os.path.join(_HERE, "port", f) for f in os.listdir(os.path.join(_HERE, "port"))]
    srcs += [os.path.join(_HERE, "shim", "gsl_shim.c"), os.path.join(_HERE, "..", "include", "tsdslam_b200.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.run(["make", "-C", _HERE, "port"], check=True, stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.port_grid_create.restype = C.c_void_p
        L.port_grid_create.argtypes = [C.c_double, C.c_int, C.c_int]
        L.port_grid_destroy.argtypes = [C.c_void_p]
        L.port_grid_set_max_truncation.argtypes = [C.c_void_p, C.c_double]
        L.port_grid_get_geometry.argtypes = [C.c_void_p, _ip, _ip, _ip] + [_dp] * 6
        L.port_grid_free_footprint.argtypes = [C.c_void_p] + [C.c_double] * 4
        L.port_grid_push.argtypes = [C.c_void_p, _sp]
        L.port_grid_last_push_stats.argtypes = [C.c_void_p, C.POINTER(PushStats)]
        L.port_grid_interpolate_bilinear.argtypes = [C.c_void_p, C.c_int32, _dp, _dp, _ip]
        L.port_grid_interpolate_normal.argtypes = [C.c_void_p, C.c_int32, _dp, _dp, _ip]
        L.port_grid_num_partitions.argtypes = [C.c_void_p]
        L.port_grid_store.argtypes = [C.c_void_p, C.c_char_p]
        L.port_grid_load.argtypes = [C.c_char_p]
        L.port_grid_load.restype = C.c_void_p
        L.port_axis_map.argtypes = [C.c_void_p, _dp, _dp, C.POINTER(C.c_uint32), C.c_void_p]
        L.port_color_image.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.port_grid_partition_states.argtypes = [C.c_void_p, _ip, _dp]
        L.port_grid_download_partition.argtypes = [C.c_void_p, C.c_int32, _dp, _dp]
        L.port_grid_upload_partition.argtypes = [C.c_void_p, C.c_int32, _dp, _dp]
        L.port_grid_fill.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
        L.port_back_project.argtypes = [_sp, C.c_int32, _dp, _ip]
        L.port_raycast_mask.argtypes = [C.c_void_p, _sp, _dp, _dp, _dp, _bp, _up]
        L.port_raycast_steps.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.port_raycast_keys.argtypes = [C.c_void_p, _sp, _dp, C.POINTER(C.c_uint64)]
        L.port_icp_create.restype = C.c_void_p
        L.port_icp_create.argtypes = [C.c_uint32, C.c_double, C.c_double, C.c_uint32, _dp]
        L.port_icp_destroy.argtypes = [C.c_void_p]
        L.port_icp_run.argtypes = [C.c_void_p, _dp, _dp, C.c_int32, _dp, C.c_int32, _dp, _dp, _dp, _dp, _up, _up, _ip]
        L.port_icp_get_trace.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _up, _up, _ip, _dp, _dp, _ip]
        L.port_match_score_tsd.argtypes = [C.c_void_p, C.c_int32, _hp, C.c_int32, _dp, _dp, _dp, _dp, C.c_double, C.c_int32,
                                           _dp, _dp, C.c_double, _dp, _ip, _dp]
        L.port_match_score_rnm.argtypes = [C.c_int32, _hp, C.c_int32, _dp, _dp, _dp, _dp, C.c_double, C.c_int32, _dp, _dp,
                                           C.c_int32, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint32,
                                           _ip, _ip, _dp, _ip, _dp]
        L.port_match_score_pdf.argtypes = [C.c_int32, _hp, C.c_int32, _dp, _dp, _dp, _dp, C.c_double, C.c_int32, _dp,
                                           C.c_int32, _dp, _dp, _dp, _dp, _ip, _ip, _dp]
        L.port_seed.argtypes = [C.c_uint32]
        L.port_match_tsd.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_uint32, C.c_double, _dp, C.c_int32, _dp, _bp,
                                     _dp, _bp, C.c_double, C.c_double, C.c_double, _dp]
        L.port_match_rnm.argtypes = [C.c_uint32, C.c_double, C.c_uint32, C.c_int32, _dp, _bp, _dp, _bp, C.c_double,
                                     C.c_double, C.c_double, _dp]
        L.port_match_pdf.argtypes = [C.c_uint32, C.c_double, C.c_uint32, _dp, C.c_int32, _dp, _bp, _dp, _bp, C.c_double,
                                     C.c_double, C.c_double, _dp]
        L.port_match_prepare.restype = C.POINTER(MatchPrepStruct)
        L.port_match_prepare.argtypes = [C.c_int32, _dp, _bp, _dp, _bp, C.c_uint32, C.c_uint32, C.c_double, C.c_double]
        L.port_match_prep_free.argtypes = [C.POINTER(MatchPrepStruct)]
        L.port_invert3x3.argtypes = [_dp, _dp]
        _lib = L
    return _lib


class MatchPrepStruct(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_control", C.c_int32), ("n_valid_m", C.c_int32), ("n_valid_s", C.c_int32),
                ("n_hyp", C.c_int32), ("span", C.c_int32), ("phi_max", C.c_double), ("theta_min", C.c_double),
                ("theta_max", C.c_double), ("phi_m", _dp), ("phi_s", _dp), ("mask_m_pca", _bp), ("mask_s_pca", _bp),
                ("idx_m_valid", _ip), ("idx_s_valid", _ip), ("idx_control", _ip), ("control", _dp), ("phi_control", _dp),
                ("hyps", _hp)]


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def invert3x3(m):
    m = _f64(m)
    out = np.empty((3, 3))
    lib().port_invert3x3(_d(m), _d(out))
    return out


def seed(s: int):
    lib().port_seed(int(s) & 0xFFFFFFFF)


def back_project(scan: Scan, xy):
    xy = _f64(xy)
    idx = np.empty(len(xy), dtype=np.int32)
    lib().port_back_project(scan.byref(), len(xy), _d(xy), idx.ctypes.data_as(_ip))
    return idx


def _file_header(path: str):
    """(cell_size, layout_partition, layout_grid) of a stored grid (first three lines of the reference's format)."""
    with open(path) as f:
        return float(f.readline()), int(f.readline()), int(f.readline())


class Grid:
    def __init__(self, cell_size: float, layout_partition: int, layout_grid: int, handle=None):
        self.h = lib().port_grid_create(cell_size, layout_partition, layout_grid) if handle is None else handle
        self.cell_size = cell_size
        self.dim = 1 << layout_partition
        self.cells = 1 << layout_grid
        self.parts_per_side = self.cells // self.dim
        self.n_partitions = self.parts_per_side ** 2

    def __del__(self):
        if getattr(self, "h", None):
            lib().port_grid_destroy(self.h)
            self.h = None

    @classmethod
    def load(cls, path: str):
        """TsdGrid(path, FILE_SOURCE)"""
        cs, lp, lg = _file_header(path)
        h = lib().port_grid_load(path.encode())
        if not h:
            raise RuntimeError(f"port_grid_load({path}) failed")
        return cls(cs, lp, lg, handle=h)

    def store(self, path: str) -> bool:
        return bool(lib().port_grid_store(self.h, path.encode()))

    def set_max_truncation(self, v):
        lib().port_grid_set_max_truncation(self.h, v)

    @property
    def bounds(self):
        v = [C.c_double() for _ in range(4)]
        lib().port_grid_get_geometry(self.h, None, None, None, None, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]),
                                     C.byref(v[3]), None)
        return tuple(x.value for x in v)

    def free_footprint(self, cx, cy, w, h) -> bool:
        return lib().port_grid_free_footprint(self.h, cx, cy, w, h) == 0

    def push(self, scan: Scan):
        lib().port_grid_push(self.h, scan.byref())

    def last_push_stats(self):
        st = PushStats()
        lib().port_grid_last_push_stats(self.h, C.byref(st))
        return st.as_dict()

    def partition_states(self):
        st = np.empty(self.n_partitions, dtype=np.int32)
        iw = np.empty(self.n_partitions)
        lib().port_grid_partition_states(self.h, st.ctypes.data_as(_ip), _d(iw))
        return st, iw

    def download_partition(self, p: int):
        n = (self.dim + 1) ** 2
        tsd = np.empty(n)
        w = np.empty(n)
        if lib().port_grid_download_partition(self.h, p, _d(tsd), _d(w)) != 0:
            return None
        return tsd.reshape(self.dim + 1, -1), w.reshape(self.dim + 1, -1)

    def upload_partition(self, p: int, tsd, w):
        tsd, w = _f64(tsd), _f64(w)
        lib().port_grid_upload_partition(self.h, p, _d(tsd), _d(w))

    def fill(self, tsd, weight, only_uninitialized: bool = False):
        lib().port_grid_fill(self.h, tsd, weight, 1 if only_uninitialized else 0)

    def interpolate_bilinear(self, xy):
        xy = _f64(xy)
        tsd = np.empty(len(xy))
        st = np.empty(len(xy), dtype=np.int32)
        lib().port_grid_interpolate_bilinear(self.h, len(xy), _d(xy), _d(tsd), st.ctypes.data_as(_ip))
        return tsd, st

    def interpolate_normal(self, xy):
        xy = _f64(xy)
        nn = np.empty((len(xy), 2))
        ok = np.empty(len(xy), dtype=np.int32)
        lib().port_grid_interpolate_normal(self.h, len(xy), _d(xy), _d(nn), ok.ctypes.data_as(_ip))
        return nn, ok

    def raycast_mask(self, scan: Scan, rays_world, coords=None, normals=None):
        n = scan.n
        rays = _f64(rays_world)
        coords = np.zeros((n, 2)) if coords is None else coords
        normals = np.zeros((n, 2)) if normals is None else normals
        mask = np.zeros(n, dtype=np.uint8)
        cnt = C.c_uint32()
        lib().port_raycast_mask(self.h, scan.byref(), _d(rays), _d(coords), _d(normals), mask.ctypes.data_as(_bp),
                                C.byref(cnt))
        return coords, normals, mask, int(cnt.value)

    def axis_map(self, with_normals: bool = False, occupied=None, cap_doubles=None):
        """RayCastAxisAligned2D::calcCoords: (coords (k, 2), normals (k, 2) or None, occupied int8[cells*cells])."""
        cap = cap_doubles or self.cells * self.cells
        coords = np.zeros(cap)
        normals = np.full(cap, np.nan) if with_normals else None
        occ = np.full(self.cells * self.cells, -1, dtype=np.int8) if occupied is None else occupied
        cnt = C.c_uint32()
        lib().port_axis_map(self.h, _d(coords), _d(normals) if with_normals else None, C.byref(cnt),
                            occ.ctypes.data_as(C.c_void_p))
        k = int(cnt.value) // 2
        return coords[:2 * k].reshape(-1, 2), (normals[:2 * k].reshape(-1, 2) if with_normals else None), occ

    def color_image(self, width: int, height: int):
        img = np.zeros(3 * width * height, dtype=np.uint8)
        lib().port_color_image(self.h, img.ctypes.data_as(C.c_void_p), width, height)
        return img.reshape(height, width, 3)

    def raycast_steps(self):
        a, b = C.c_uint64(), C.c_uint64()
        lib().port_raycast_steps(C.byref(a), C.byref(b))
        return a.value, b.value

    def raycast_keys(self, scan: Scan, rays_world):
        rays = _f64(rays_world)
        keys = np.empty(scan.n, dtype=np.uint64)
        lib().port_raycast_keys(self.h, scan.byref(), _d(rays), keys.ctypes.data_as(C.POINTER(C.c_uint64)))
        return keys


class Icp:
    def __init__(self, max_iterations, dist_max, dist_min, bounds):
        b = _f64(bounds)
        self.max_iterations = max_iterations
        self.h = lib().port_icp_create(max_iterations, dist_max, dist_min, (max_iterations - 10) & 0xFFFFFFFF, _d(b))

    def __del__(self):
        if getattr(self, "h", None):
            lib().port_icp_destroy(self.h)
            self.h = None

    def run(self, model, normals, scene, pose, Tinit44=None):
        model, normals, scene, pose = _f64(model), _f64(normals), _f64(scene), _f64(pose)
        Ti = None if Tinit44 is None else _f64(Tinit44)
        T = np.empty((3, 3))
        mse, pairs, its, st = C.c_double(), C.c_uint32(), C.c_uint32(), C.c_int32()
        lib().port_icp_run(self.h, _d(model), _d(normals), len(model), _d(scene), len(scene), _d(pose),
                           None if Ti is None else _d(Ti), _d(T), C.byref(mse), C.byref(pairs), C.byref(its), C.byref(st))
        return T, mse.value, pairs.value, its.value, st.value

    def set_trace(self, enable: bool = True):
        pass  # the port always records

    def trace(self, cap):
        mi = self.max_iterations
        pm = np.zeros((mi, cap), dtype=np.uint32)
        ps = np.zeros((mi, cap), dtype=np.uint32)
        pc = np.zeros(mi, dtype=np.int32)
        mse = np.zeros(mi)
        Tf = np.zeros((mi, 4, 4))
        nit = C.c_int32()
        lib().port_icp_get_trace(self.h, mi, cap, pm.ctypes.data_as(_up), ps.ctypes.data_as(_up), pc.ctypes.data_as(_ip),
                                 _d(mse), _d(Tf), C.byref(nit))
        return nit.value, pm, ps, pc, mse, Tf


class MatchPrep:
    """Numpy view of port_match_prepare's output (copied)."""

    def __init__(self, n, M, maskM, S, maskS, trials, size_control, phi_max, resolution):
        M, S = _f64(M), _f64(S)
        mM, mS = _u8(maskM), _u8(maskS)
        p = lib().port_match_prepare(n, _d(M), mM.ctypes.data_as(_bp), _d(S), mS.ctypes.data_as(_bp), trials, size_control,
                                     phi_max, resolution)
        self.ok = bool(p)
        if not self.ok:
            return
        c = p.contents
        self.n = c.n
        self.n_control, self.n_valid_m, self.n_valid_s, self.n_hyp, self.span = (c.n_control, c.n_valid_m, c.n_valid_s,
                                                                                   c.n_hyp, c.span)
        self.phi_max, self.theta_min, self.theta_max = c.phi_max, c.theta_min, c.theta_max
        arr = np.ctypeslib.as_array
        self.phi_m = arr(c.phi_m, (c.n,)).copy()
        self.phi_s = arr(c.phi_s, (c.n,)).copy()
        self.mask_m_pca = arr(c.mask_m_pca, (c.n,)).copy()
        self.mask_s_pca = arr(c.mask_s_pca, (c.n,)).copy()
        self.idx_m_valid = arr(c.idx_m_valid, (c.n_valid_m,)).copy()
        self.idx_s_valid = arr(c.idx_s_valid, (c.n_valid_s,)).copy()
        self.idx_control = arr(c.idx_control, (max(c.n_control, 1),)).copy()[:c.n_control]
        self.control = arr(c.control, (3 * max(c.n_control, 1),)).copy()[:3 * c.n_control].reshape(3, c.n_control)
        self.phi_control = arr(c.phi_control, (max(c.n_control, 1),)).copy()[:c.n_control]
        hy = np.ctypeslib.as_array(C.cast(c.hyps, _ip), (max(c.n_hyp, 1) * 2,)).copy()[:2 * c.n_hyp]
        self.hyps = hy.reshape(-1, 2).astype(np.int32)
        lib().port_match_prep_free(p)


def _hyps(h):
    h = np.ascontiguousarray(h, dtype=np.int32).reshape(-1, 2)
    return h, h.ctypes.data_as(_hp)


def score_tsd(grid: Grid, hyps, M, S, phi_m, phi_s, phi_max, control, t_sensor, zrand):
    h, hp = _hyps(hyps)
    M, S, phi_m, phi_s, control, t_sensor = map(_f64, (M, S, phi_m, phi_s, control, t_sensor))
    score = np.empty(len(h))
    best = C.c_int32()
    T = np.empty((3, 3))
    lib().port_match_score_tsd(grid.h, len(h), hp, len(M), _d(M), _d(S), _d(phi_m), _d(phi_s), phi_max, control.shape[1],
                               _d(control), _d(t_sensor), zrand, _d(score), C.byref(best), _d(T))
    return score, best.value, T


def score_rnm(hyps, M, S, phi_m, phi_s, phi_max, control, phi_control, model_valid, phi_valid, theta_min, theta_max,
              scale_distance, scale_orientation, cnt_thresh):
    h, hp = _hyps(hyps)
    M, S, phi_m, phi_s, control, phi_control, model_valid, phi_valid = map(
        _f64, (M, S, phi_m, phi_s, control, phi_control, model_valid, phi_valid))
    cnt = np.empty(len(h), dtype=np.int32)
    mx = np.empty(len(h), dtype=np.int32)
    err = np.empty(len(h))
    best = C.c_int32()
    T = np.empty((3, 3))
    lib().port_match_score_rnm(len(h), hp, len(M), _d(M), _d(S), _d(phi_m), _d(phi_s), phi_max, control.shape[1],
                               _d(control), _d(phi_control), len(model_valid), _d(model_valid), _d(phi_valid), theta_min,
                               theta_max, scale_distance, scale_orientation, cnt_thresh, cnt.ctypes.data_as(_ip),
                               mx.ctypes.data_as(_ip), _d(err), C.byref(best), _d(T))
    return cnt, mx, err, best.value, T


def score_pdf(hyps, M, S, phi_m, phi_s, phi_max, control, model_angles, model_dists, params):
    h, hp = _hyps(hyps)
    M, S, phi_m, phi_s, control, model_angles, model_dists, params = map(
        _f64, (M, S, phi_m, phi_s, control, model_angles, model_dists, params))
    prob = np.empty(len(h))
    fov = np.empty(len(h), dtype=np.int32)
    best = C.c_int32()
    T = np.empty((3, 3))
    lib().port_match_score_pdf(len(h), hp, len(M), _d(M), _d(S), _d(phi_m), _d(phi_s), phi_max, control.shape[1],
                               _d(control), len(model_angles), _d(model_angles), _d(model_dists), _d(params), _d(prob),
                               fov.ctypes.data_as(_ip), C.byref(best), _d(T))
    return prob, fov, best.value, T


def match_tsd(grid: Grid, trials, eps, size_control, zrand, TSensor, M, maskM, S, maskS, phi_max, trans_max, resolution):
    M, S, TSensor = _f64(M), _f64(S), _f64(TSensor)
    mM, mS = _u8(maskM), _u8(maskS)
    T = np.empty((3, 3))
    lib().port_match_tsd(grid.h, trials, eps, size_control, zrand, _d(TSensor), len(M), _d(M), mM.ctypes.data_as(_bp), _d(S),
                         mS.ctypes.data_as(_bp), phi_max, trans_max, resolution, _d(T))
    return T


def match_rnm(trials, eps, size_control, M, maskM, S, maskS, phi_max, trans_max, resolution):
    M, S = _f64(M), _f64(S)
    mM, mS = _u8(maskM), _u8(maskS)
    T = np.empty((3, 3))
    lib().port_match_rnm(trials, eps, size_control, len(M), _d(M), mM.ctypes.data_as(_bp), _d(S), mS.ctypes.data_as(_bp),
                         phi_max, trans_max, resolution, _d(T))
    return T


def match_pdf(trials, eps, size_control, params, M, maskM, S, maskS, phi_max, trans_max, resolution):
    M, S, params = _f64(M), _f64(S), _f64(params)
    mM, mS = _u8(maskM), _u8(maskS)
    T = np.empty((3, 3))
    lib().port_match_pdf(trials, eps, size_control, _d(params), len(M), _d(M), mM.ctypes.data_as(_bp), _d(S),
                         mS.ctypes.data_as(_bp), phi_max, trans_max, resolution, _d(T))
    return T
