"""ctypes mirror of `tsd_scan_t` (include/tsdslam_b200.h) and a host-side sensor model.

`HostSensor` keeps what obvious::SensorPolar2D keeps on the host (pose, ray map, measurement data and
mask; reference src/obvision/reconstruct/Sensor.cpp, grid/SensorPolar2D.cpp) -- this part of the
reference stays on the host; the kernels only read the POD scan.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np


class ScanStruct(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("_pad", C.c_int32),
        ("ranges", C.POINTER(C.c_double)),
        ("mask", C.POINTER(C.c_ubyte)),
        ("pose", C.c_double * 9),
        ("pose_inv", C.c_double * 9),
        ("phi_min", C.c_double),
        ("angular_res", C.c_double),
        ("phi_lower", C.c_double),
        ("phi_upper", C.c_double),
        ("max_range", C.c_double),
        ("min_range", C.c_double),
        ("low_reflectivity_range", C.c_double),
    ]


class PushStats(C.Structure):
    _fields_ = [
        ("cell_updates", C.c_uint64),
        ("cell_visits", C.c_uint64),
        ("active_tiles", C.c_uint32),
        ("emptied_tiles", C.c_uint32),
        ("newly_initialized", C.c_uint32),
        ("fallback_cells", C.c_uint32),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Hypothesis(C.Structure):
    _fields_ = [("idx_model", C.c_int32), ("idx_scene", C.c_int32)]


class Scan:
    """One scan as the kernels see it.  Keeps the numpy arrays alive for the ctypes struct."""

    def __init__(self, spec, ranges: np.ndarray, mask: np.ndarray, pose: np.ndarray, pose_inv: np.ndarray,
                 phi_lower: float | None = None, phi_upper: float | None = None):
        self.spec = spec
        self.ranges = np.ascontiguousarray(ranges, dtype=np.float64)
        self.mask = np.ascontiguousarray(mask, dtype=np.uint8)
        self.pose = np.ascontiguousarray(pose, dtype=np.float64).reshape(3, 3)
        self.pose_inv = np.ascontiguousarray(pose_inv, dtype=np.float64).reshape(3, 3)
        s = ScanStruct()
        s.n = len(self.ranges)
        s.ranges = self.ranges.ctypes.data_as(C.POINTER(C.c_double))
        s.mask = self.mask.ctypes.data_as(C.POINTER(C.c_ubyte))
        for i in range(9):
            s.pose[i] = float(self.pose.flat[i])
            s.pose_inv[i] = float(self.pose_inv.flat[i])
        s.phi_min = spec.phi_min
        s.angular_res = spec.angular_res
        s.phi_lower = spec.phi_lower if phi_lower is None else phi_lower
        s.phi_upper = spec.phi_upper if phi_upper is None else phi_upper
        s.max_range = spec.max_range
        s.min_range = spec.min_range
        s.low_reflectivity_range = spec.low_reflectivity_range
        self.struct = s

    @property
    def n(self):
        return len(self.ranges)

    def byref(self):
        return C.byref(self.struct)


def standard_mask(ranges64: np.ndarray, spec) -> tuple[np.ndarray, np.ndarray]:
    """SensorPolar2D::setStandardMask (SensorPolar2D.cpp:59-98; Sensor.cpp:246-272), host side.

    Returns (data, mask); like the reference, data above max range or NaN becomes +inf."""
    data = np.array(ranges64, dtype=np.float64, copy=True)
    n = len(data)
    mask = np.ones(n, dtype=np.uint8)
    mask &= (data != 0.0).astype(np.uint8)                       # maskZeroDepth
    data[data > spec.max_range] = np.inf                          # maskInvalidDepth
    nan = np.isnan(data)
    mask[nan] = 0
    data[nan] = np.inf
    # maskDepthDiscontinuity(deg2rad(3.0)), radius 1
    thresh = (math.pi * 3.0) / 180.0
    sinphi = math.sin(spec.angular_res)
    cosphi = math.cos(spec.angular_res)
    for i in range(1, n - 1):
        betamin = math.pi
        a = data[i]
        if math.isinf(a):
            continue
        for j in (-1, 0, 1):
            b = data[i + j]
            if math.isinf(b):
                continue
            c2 = a * a + b * b - 2 * a * b * cosphi
            c = math.sqrt(c2) if c2 >= 0 else float("nan")
            if a > b:
                with np.errstate(all="ignore"):
                    q = b / c * sinphi if c != 0 else float("inf")
                beta = math.asin(q) if -1.0 <= q <= 1.0 else float("nan")
                if beta < betamin:
                    betamin = beta
        if betamin < thresh:
            mask[i] = 0
    return data, mask


def gemm_nn(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """A @ B with the rounding order of gslcblas dgemm NoTrans x NoTrans (row major): k outer, zero
    coefficients skipped (used by operator* in obcore/math/linalg/gsl/Matrix.cpp:90-95)."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    Cm = np.zeros((A.shape[0], B.shape[1]))
    for k in range(A.shape[1]):
        t = A[:, k][:, None]
        with np.errstate(invalid="ignore"):
            Cm = np.where(t != 0.0, Cm + t * B[k, :][None, :], Cm)
    return Cm


class HostSensor:
    """Host-side state of obvious::SensorPolar2D (pose, ray map, data, mask) in numpy.

    Mirrors SensorPolar2D.cpp:11-48 (ctor), Sensor.cpp:36-60 (getNormalizedRayMap, transform),
    Sensor.cpp:125-145 (setRealMeasurementData), SensorPolar2D.cpp:59-65 (setStandardMask) and
    Sensor.cpp:168-190 (dataToCartesianVectorMask).  `invert` is the 3x3 inverse routine to use for
    pose_inv (the C-ABI's tsd_invert3x3, or the oracle's when pinning)."""

    def __init__(self, spec, invert):
        self.spec = spec
        self.n = spec.beams
        self._invert = invert
        phis = [spec.phi_min + float(i) * spec.angular_res for i in range(self.n)]
        self.rays = np.array([[math.cos(p) for p in phis], [math.sin(p) for p in phis]], dtype=np.float64)
        self.rays_local = self.rays.copy()
        self.ray_norm = 1.0
        self.T = np.eye(3)
        self.data = np.zeros(self.n)
        self.mask = np.ones(self.n, dtype=np.uint8)

    def set_scan(self, ranges_f32, standard_mask_: bool = True):
        r = np.asarray(ranges_f32, dtype=np.float32)
        self.data = (r * np.float32(1.0)).astype(np.float64)  # Sensor.cpp:143-144: (double)(data[i] * scale)
        if standard_mask_:
            self.data, self.mask = standard_mask(self.data, self.spec)

    def transform(self, T):
        T = np.asarray(T, dtype=np.float64).reshape(3, 3)
        self.rays = gemm_nn(T[:2, :2], self.rays)  # Sensor.cpp:52-54
        self.T = gemm_nn(self.T, T)                # Sensor.cpp:59

    def normalized_rays(self, norm: float) -> np.ndarray:
        if norm != self.ray_norm:                  # Sensor.cpp:38-46
            self.rays = self.rays * (norm / self.ray_norm)
            self.ray_norm = norm
        return self.rays

    @property
    def pose(self):
        return self.T

    def scan(self) -> Scan:
        return Scan(self.spec, self.data, self.mask, self.T, self._invert(self.T))

    def scene(self):
        """dataToCartesianVectorMask: (n x 2 coords in the sensor frame, mask, count)."""
        valid = (~np.isinf(self.data)) & (self.mask != 0)
        coords = np.zeros((self.n, 2))
        coords[valid, 0] = self.rays_local[0, valid] * self.data[valid]
        coords[valid, 1] = self.rays_local[1, valid] * self.data[valid]
        return coords, valid.astype(np.uint8), int(valid.sum())
