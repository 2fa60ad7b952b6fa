"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libohm_ref.so (oracle/ref_capi.cpp).

The library is the reference's own C++ (obvious::TsdGrid, SensorPolar2D, RayCastPolar2D, Icp, matchers)
compiled unmodified against the GSL/FLANN shims.  It is built in the authoring container (where
/root/reference exists) and travels to the GPU box as a prebuilt file.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libohm_ref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint)
_bp = C.POINTER(C.c_ubyte)


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        L.ref_grid_create.restype = C.c_void_p
        L.ref_grid_create.argtypes = [C.c_double, C.c_int, C.c_int]
        L.ref_grid_load.restype = C.c_void_p
        L.ref_grid_load.argtypes = [C.c_char_p]
        L.ref_grid_store.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_grid_destroy.argtypes = [C.c_void_p]
        L.ref_axis_map.argtypes = [C.c_void_p, _dp, _dp, C.c_void_p]
        L.ref_axis_map.restype = C.c_uint
        L.ref_color_image.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint]
        L.ref_grid_set_max_truncation.argtypes = [C.c_void_p, C.c_double]
        for f in ("ref_grid_get_max_truncation", "ref_grid_min_x", "ref_grid_max_x", "ref_grid_min_y", "ref_grid_max_y"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_grid_cells_x.argtypes = [C.c_void_p]
        L.ref_grid_free_footprint.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
        L.ref_grid_push.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_grid_fill.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
        L.ref_grid_num_partitions.argtypes = [C.c_void_p]
        L.ref_grid_partition_state.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_grid_partition_states.argtypes = [C.c_void_p, _ip, _dp]
        L.ref_grid_download_partition.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.ref_grid_interpolate_bilinear.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _ip]
        L.ref_grid_interpolate_normal.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _ip]
        L.ref_sensor_create.restype = C.c_void_p
        L.ref_sensor_create.argtypes = [C.c_int] + [C.c_double] * 5
        L.ref_sensor_destroy.argtypes = [C.c_void_p]
        L.ref_sensor_set_data.argtypes = [C.c_void_p, _dp]
        L.ref_sensor_set_data_f32.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
        L.ref_sensor_set_standard_mask.argtypes = [C.c_void_p]
        L.ref_sensor_set_mask.argtypes = [C.c_void_p, _bp]
        L.ref_sensor_get_data.argtypes = [C.c_void_p, _dp]
        L.ref_sensor_get_mask.argtypes = [C.c_void_p, _bp]
        L.ref_sensor_transform.argtypes = [C.c_void_p, _dp]
        L.ref_sensor_set_pose.argtypes = [C.c_void_p, _dp]
        L.ref_sensor_get_pose.argtypes = [C.c_void_p, _dp]
        L.ref_sensor_get_normalized_rays.argtypes = [C.c_void_p, C.c_double, _dp]
        L.ref_sensor_phi_lower.restype = C.c_double
        L.ref_sensor_phi_lower.argtypes = [C.c_void_p]
        L.ref_sensor_phi_upper.restype = C.c_double
        L.ref_sensor_phi_upper.argtypes = [C.c_void_p]
        L.ref_sensor_back_project.argtypes = [C.c_void_p, C.c_int, _dp, _ip]
        L.ref_sensor_data_to_cartesian_mask.restype = C.c_uint
        L.ref_sensor_data_to_cartesian_mask.argtypes = [C.c_void_p, _dp, _bp]
        L.ref_raycast_mask.restype = C.c_uint
        L.ref_raycast_mask.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, _bp]
        L.ref_raycast_compact.restype = C.c_uint
        L.ref_raycast_compact.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp]
        L.ref_icp_create.restype = C.c_void_p
        L.ref_icp_create.argtypes = [C.c_uint, C.c_double, C.c_double, C.c_uint] + [C.c_double] * 4
        L.ref_icp_destroy.argtypes = [C.c_void_p]
        L.ref_icp_run.argtypes = [C.c_void_p, _dp, _dp, C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, _up, _up]
        L.ref_icp_trace.argtypes = [C.c_void_p, _dp, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, _up, _up, _ip, _dp, _dp]
        L.ref_match_tsd.argtypes = [C.c_void_p, C.c_uint, C.c_double, C.c_uint, C.c_double, _dp, C.c_int, _dp, _bp, _dp, _bp,
                                    C.c_double, C.c_double, C.c_double, _dp]
        L.ref_match_rnm.argtypes = [C.c_uint, C.c_double, C.c_uint, C.c_int, _dp, _bp, _dp, _bp, C.c_double, C.c_double,
                                    C.c_double, _dp]
        L.ref_match_pdf.argtypes = [C.c_uint, C.c_double, C.c_uint, _dp, C.c_int, _dp, _bp, _dp, _bp, C.c_double, C.c_double,
                                    C.c_double, _dp]
        L.ref_pdf_probability.restype = C.c_double
        L.ref_pdf_probability.argtypes = [_dp, C.c_double, C.c_double, C.c_double]
        L.ref_invert.argtypes = [C.c_int, _dp, _dp]
        L.ref_matmul.argtypes = [C.c_int, _dp, _dp, _dp]
        L.ref_seed.argtypes = [C.c_uint]
        L.ref_rand_calls.restype = C.c_ulonglong
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_quiet()
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def invert(m: np.ndarray) -> np.ndarray:
    m = _f64(m)
    out = np.empty_like(m)
    lib().ref_invert(m.shape[0], _d(m), _d(out))
    return out


def matmul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a, b = _f64(a), _f64(b)
    out = np.empty_like(a)
    lib().ref_matmul(a.shape[0], _d(a), _d(b), _d(out))
    return out


def seed(s: int):
    lib().ref_seed(int(s) & 0xFFFFFFFF)


def set_threads(n: int):
    lib().ref_set_threads(int(n))


def max_threads() -> int:
    return lib().ref_max_threads()


class Sensor:
    """obvious::SensorPolar2D"""

    def __init__(self, spec):
        self.spec = spec
        self.n = spec.beams
        self.h = lib().ref_sensor_create(spec.beams, spec.angular_res, spec.phi_min, spec.max_range, spec.min_range,
                                         spec.low_reflectivity_range)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_sensor_destroy(self.h)
            self.h = None

    def set_scan(self, ranges_f32: np.ndarray, standard_mask: bool = True):
        r = np.ascontiguousarray(ranges_f32, dtype=np.float32)
        lib().ref_sensor_set_data_f32(self.h, r.ctypes.data_as(C.POINTER(C.c_float)), len(r))
        if standard_mask:
            lib().ref_sensor_set_standard_mask(self.h)

    def set_data(self, ranges_f64, mask=None):
        r = _f64(ranges_f64)
        lib().ref_sensor_set_data(self.h, _d(r))
        if mask is not None:
            m = _u8(mask)
            lib().ref_sensor_set_mask(self.h, m.ctypes.data_as(_bp))

    @property
    def data(self) -> np.ndarray:
        out = np.empty(self.n)
        lib().ref_sensor_get_data(self.h, _d(out))
        return out

    @property
    def mask(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.uint8)
        lib().ref_sensor_get_mask(self.h, out.ctypes.data_as(_bp))
        return out

    def transform(self, T: np.ndarray):
        T = _f64(T)
        lib().ref_sensor_transform(self.h, _d(T))

    @property
    def pose(self) -> np.ndarray:
        out = np.empty((3, 3))
        lib().ref_sensor_get_pose(self.h, _d(out))
        return out

    @pose.setter
    def pose(self, T):
        T = _f64(T)
        lib().ref_sensor_set_pose(self.h, _d(T))

    def normalized_rays(self, norm: float) -> np.ndarray:
        out = np.empty((2, self.n))
        lib().ref_sensor_get_normalized_rays(self.h, norm, _d(out))
        return out

    @property
    def phi_bounds(self):
        return lib().ref_sensor_phi_lower(self.h), lib().ref_sensor_phi_upper(self.h)

    def back_project(self, xy: np.ndarray) -> np.ndarray:
        xy = _f64(xy)
        idx = np.empty(len(xy), dtype=np.int32)
        lib().ref_sensor_back_project(self.h, len(xy), _d(xy), idx.ctypes.data_as(_ip))
        return idx

    def scene(self):
        coords = np.zeros((self.n, 2))
        mask = np.zeros(self.n, dtype=np.uint8)
        valid = lib().ref_sensor_data_to_cartesian_mask(self.h, _d(coords), mask.ctypes.data_as(_bp))
        return coords, mask, int(valid)


class _quiet_stdout:
    """TsdGrid::init prints "init" with std::cout (TsdGrid.cpp:114): keep it off the caller's stdout."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


class Grid:
    """obvious::TsdGrid"""

    def __init__(self, cell_size: float, layout_partition: int, layout_grid: int, handle=None):
        if handle is None:
            L = lib()
            with _quiet_stdout():
                handle = L.ref_grid_create(cell_size, layout_partition, layout_grid)
        self.h = handle
        self.cell_size = cell_size
        self.dim = 1 << layout_partition
        self.cells = 1 << layout_grid
        self.parts_per_side = self.cells // self.dim
        self.n_partitions = self.parts_per_side ** 2

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_grid_destroy(self.h)
            self.h = None

    def set_max_truncation(self, v: float):
        lib().ref_grid_set_max_truncation(self.h, v)

    @property
    def bounds(self):
        L = lib()
        return L.ref_grid_min_x(self.h), L.ref_grid_max_x(self.h), L.ref_grid_min_y(self.h), L.ref_grid_max_y(self.h)

    def free_footprint(self, cx, cy, w, h) -> bool:
        return bool(lib().ref_grid_free_footprint(self.h, cx, cy, w, h))

    def push(self, sensor: Sensor):
        lib().ref_grid_push(self.h, sensor.h)

    def fill(self, tsd: float, weight: float, only_uninitialized: bool = False):
        lib().ref_grid_fill(self.h, tsd, weight, 1 if only_uninitialized else 0)

    def partition_states(self):
        st = np.empty(self.n_partitions, dtype=np.int32)
        iw = np.empty(self.n_partitions)
        lib().ref_grid_partition_states(self.h, st.ctypes.data_as(_ip), _d(iw))
        return st, iw

    def download_partition(self, p: int):
        n = (self.dim + 1) ** 2
        tsd = np.empty(n)
        w = np.empty(n)
        ok = lib().ref_grid_download_partition(self.h, p, _d(tsd), _d(w))
        if not ok:
            return None
        return tsd.reshape(self.dim + 1, self.dim + 1), w.reshape(self.dim + 1, self.dim + 1)

    def interpolate_bilinear(self, xy: np.ndarray):
        xy = _f64(xy)
        tsd = np.empty(len(xy))
        st = np.empty(len(xy), dtype=np.int32)
        lib().ref_grid_interpolate_bilinear(self.h, len(xy), _d(xy), _d(tsd), st.ctypes.data_as(_ip))
        return tsd, st

    def interpolate_normal(self, xy: np.ndarray):
        xy = _f64(xy)
        nn = np.empty((len(xy), 2))
        ok = np.empty(len(xy), dtype=np.int32)
        lib().ref_grid_interpolate_normal(self.h, len(xy), _d(xy), _d(nn), ok.ctypes.data_as(_ip))
        return nn, ok

    def raycast_mask(self, sensor: Sensor, coords=None, normals=None):
        n = sensor.n
        coords = np.zeros((n, 2)) if coords is None else coords
        normals = np.zeros((n, 2)) if normals is None else normals
        mask = np.zeros(n, dtype=np.uint8)
        cnt = lib().ref_raycast_mask(self.h, sensor.h, _d(coords), _d(normals), mask.ctypes.data_as(_bp))
        return coords, normals, mask, int(cnt)

    def axis_map(self, with_normals: bool = False, occupied=None, cap_doubles=None):
        """RayCastAxisAligned2D::calcCoords: (coords (k, 2), normals (k, 2) or None, occupied int8[cells*cells])."""
        cells = int(lib().ref_grid_cells_x(self.h))
        cap = cap_doubles or cells * cells
        coords = np.zeros(cap)
        normals = np.full(cap, np.nan) if with_normals else None
        occ = np.full(cells * cells, -1, dtype=np.int8) if occupied is None else occupied
        lib().ref_axis_map.restype = C.c_uint
        cnt = lib().ref_axis_map(self.h, _d(coords), _d(normals) if with_normals else None,
                                 occ.ctypes.data_as(C.c_void_p))
        k = int(cnt) // 2
        return coords[:2 * k].reshape(-1, 2), (normals[:2 * k].reshape(-1, 2) if with_normals else None), occ

    def color_image(self, width: int, height: int):
        img = np.zeros(3 * width * height, dtype=np.uint8)
        lib().ref_color_image(self.h, img.ctypes.data_as(C.c_void_p), width, height)
        return img.reshape(height, width, 3)

    def store(self, path: str) -> bool:
        return bool(lib().ref_grid_store(self.h, path.encode()))

    @classmethod
    def load(cls, path: str):
        """TsdGrid(path, FILE_SOURCE)"""
        with open(path) as f:
            cs, lp, lg = float(f.readline()), int(f.readline()), int(f.readline())
        with _quiet_stdout():
            h = lib().ref_grid_load(path.encode())
        return cls(cs, lp, lg, handle=h)


class Icp:
    """obvious::Icp wired as ThreadLocalize.cpp:210-225."""

    def __init__(self, max_iterations: int, dist_max: float, dist_min: float, bounds):
        self.max_iterations = max_iterations
        self.h = lib().ref_icp_create(max_iterations, dist_max, dist_min, (max_iterations - 10) & 0xFFFFFFFF, *bounds)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_icp_destroy(self.h)
            self.h = None

    def run(self, model, normals, scene, pose, Tinit44=None):
        model, normals, scene, pose = _f64(model), _f64(normals), _f64(scene), _f64(pose)
        Tinit = _f64(np.eye(4) if Tinit44 is None else Tinit44)
        T = np.empty((3, 3))
        rms = C.c_double()
        pairs = C.c_uint()
        its = C.c_uint()
        st = lib().ref_icp_run(self.h, _d(model), _d(normals), len(model), _d(scene), len(scene), _d(pose), _d(Tinit), _d(T),
                               C.byref(rms), C.byref(pairs), C.byref(its))
        return T, rms.value, pairs.value, its.value, st

    def trace(self, model, normals, scene, pose):
        model, normals, scene, pose = _f64(model), _f64(normals), _f64(scene), _f64(pose)
        cap = max(len(model), len(scene))
        mi = self.max_iterations
        pm = np.zeros((mi, cap), dtype=np.uint32)
        ps = np.zeros((mi, cap), dtype=np.uint32)
        pc = np.zeros(mi, dtype=np.int32)
        rms = np.zeros(mi)
        Tf = np.zeros((mi, 4, 4))
        its = lib().ref_icp_trace(self.h, _d(model), _d(normals), len(model), _d(scene), len(scene), _d(pose), mi, cap,
                                  pm.ctypes.data_as(_up), ps.ctypes.data_as(_up), pc.ctypes.data_as(_ip), _d(rms), _d(Tf))
        return its, pm, ps, pc, rms, Tf


PDF_DEFAULTS = np.array([0.45, 0.0, 0.25, 0.05, 0.25, 0.9, 20.0, np.pi / 180.0 * 3, 0.2, 0.08, 3.0, 0.5])


def match_tsd(grid: Grid, trials, eps, size_control, zrand, TSensor, M, maskM, S, maskS, phi_max, trans_max, resolution):
    M, S, TSensor = _f64(M), _f64(S), _f64(TSensor)
    mM, mS = _u8(maskM), _u8(maskS)
    T = np.empty((3, 3))
    lib().ref_match_tsd(grid.h, trials, eps, size_control, zrand, _d(TSensor), len(M), _d(M), mM.ctypes.data_as(_bp), _d(S),
                        mS.ctypes.data_as(_bp), phi_max, trans_max, resolution, _d(T))
    return T


def match_rnm(trials, eps, size_control, M, maskM, S, maskS, phi_max, trans_max, resolution):
    M, S = _f64(M), _f64(S)
    mM, mS = _u8(maskM), _u8(maskS)
    T = np.empty((3, 3))
    lib().ref_match_rnm(trials, eps, size_control, len(M), _d(M), mM.ctypes.data_as(_bp), _d(S), mS.ctypes.data_as(_bp),
                        phi_max, trans_max, resolution, _d(T))
    return T


def match_pdf(trials, eps, size_control, params, M, maskM, S, maskS, phi_max, trans_max, resolution):
    M, S, params = _f64(M), _f64(S), _f64(params)
    mM, mS = _u8(maskM), _u8(maskS)
    T = np.empty((3, 3))
    lib().ref_match_pdf(trials, eps, size_control, _d(params), len(M), _d(M), mM.ctypes.data_as(_bp), _d(S),
                        mS.ctypes.data_as(_bp), phi_max, trans_max, resolution, _d(T))
    return T
