"""Synthetic Hokuyo-like scans: the only input data of tests and bench.

World: an axis-aligned rectangular room centred in the grid plus K random convex
polygon obstacles; analytic ray casting in FP64.  Sensor: N beams, 270 deg FOV,
Gaussian range noise, a fraction of +inf (no return) and 0 (masked by
maskZeroDepth) beams; ranges are quantised to float32 and widened again, as the
node does (reference src/obvision/reconstruct/Sensor.cpp:136-145 takes vector<float>).

All of it is plain numpy so that it runs on the GPU box (no /root/reference there).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

BASE_SEED = 20260101


@dataclass
class SensorSpec:
    """Constructor arguments of obvious::SensorPolar2D (SensorPolar2D.cpp:11)."""

    beams: int = 1081
    angular_res: float = math.pi / 720.0  # 0.25 deg
    phi_min: float = -135.0 * math.pi / 180.0
    max_range: float = 30.0
    min_range: float = 0.001
    low_reflectivity_range: float = 2.0

    @property
    def phi_lower(self) -> float:  # SensorPolar2D.cpp:26
        return -0.5 * self.angular_res + self.phi_min

    @property
    def phi_upper(self) -> float:  # SensorPolar2D.cpp:30
        return self.phi_min + (float(self.beams) - 0.5) * self.angular_res


@dataclass
class World:
    segments: np.ndarray  # (S, 4) x0 y0 x1 y1

    @staticmethod
    def room_with_obstacles(cx: float, cy: float, width: float, height: float, n_obstacles: int, seed: int,
                            keep_clear: float = 1.0, obstacle_scale: float | None = None) -> "World":
        """obstacle_scale: circumradius of the obstacles in units of 0.15 .. 0.6 m; default: grows with the room
        (1 for rooms up to 20 m)."""
        rng = np.random.default_rng(seed)
        x0, x1 = cx - width / 2, cx + width / 2
        y0, y1 = cy - height / 2, cy + height / 2
        segs = [(x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)]
        made = 0
        tries = 0
        while made < n_obstacles and tries < 100 * max(n_obstacles, 1):
            tries += 1
            r = rng.uniform(0.15, 0.6) * (max(1.0, min(width, height) / 20.0) if obstacle_scale is None else obstacle_scale)
            ox = rng.uniform(x0 + r, x1 - r)
            oy = rng.uniform(y0 + r, y1 - r)
            if math.hypot(ox - cx, oy - cy) < keep_clear + r + 0.5:
                continue
            k = int(rng.integers(3, 7))
            ang = np.sort(rng.uniform(0, 2 * math.pi, k))
            if np.any(np.diff(np.concatenate([ang, ang[:1] + 2 * math.pi])) > math.pi * 0.95):
                continue
            px = ox + r * np.cos(ang)
            py = oy + r * np.sin(ang)
            for i in range(k):
                j = (i + 1) % k
                segs.append((px[i], py[i], px[j], py[j]))
            made += 1
        return World(np.asarray(segs, dtype=np.float64))

    def cast(self, px: float, py: float, angles: np.ndarray) -> np.ndarray:
        """Exact ranges along `angles` (world frame) from (px, py); +inf where nothing is hit."""
        dx = np.cos(angles)[:, None]
        dy = np.sin(angles)[:, None]
        ax = self.segments[None, :, 0] - px
        ay = self.segments[None, :, 1] - py
        ex = self.segments[None, :, 2] - self.segments[None, :, 0]
        ey = self.segments[None, :, 3] - self.segments[None, :, 1]
        den = dx * ey - dy * ex
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (ax * ey - ay * ex) / den
            u = (ax * dy - ay * dx) / den
        ok = (np.abs(den) > 1e-12) & (t > 1e-9) & (u >= 0.0) & (u <= 1.0)
        t = np.where(ok, t, np.inf)
        return t.min(axis=1)


def pose_matrix(x: float, y: float, theta: float) -> np.ndarray:
    """3x3 homogeneous pose, laid out as ThreadLocalize.cpp:469-471."""
    c, s = math.cos(theta), math.sin(theta)
    return np.array([[c, -s, x], [s, c, y], [0.0, 0.0, 1.0]], dtype=np.float64)


def scan_from_pose(world: World, spec: SensorSpec, x: float, y: float, theta: float, rng: np.random.Generator,
                   noise_sigma: float = 0.01, p_inf: float = 0.01, p_zero: float = 0.005) -> np.ndarray:
    """float32 ranges as a LaserScan message would carry them."""
    ang = theta + spec.phi_min + np.arange(spec.beams, dtype=np.float64) * spec.angular_res
    r = world.cast(x, y, ang)
    r = r + rng.normal(0.0, noise_sigma, size=r.shape)
    r = np.where(r > spec.max_range, np.inf, r)
    u = rng.uniform(size=r.shape)
    r = np.where(u < p_inf, np.inf, r)
    r = np.where((u >= p_inf) & (u < p_inf + p_zero), 0.0, r)
    r = np.maximum(r, 0.0)
    return r.astype(np.float32)


@dataclass
class Config:
    """One BASELINE.json configuration."""

    name: str
    layout_grid: int
    cell_size: float = 0.025
    layout_partition: int = 5
    truncation_cells: float = 3.0
    sensor: SensorSpec = field(default_factory=SensorSpec)
    room: tuple = (20.0, 15.0)
    n_obstacles: int = 8
    seed: int = BASE_SEED
    n_scans: int = 200
    twist: tuple = (0.03, 0.0, 0.005)  # forward, lateral, yaw per scan
    obstacle_scale: float | None = None

    @property
    def cells(self) -> int:
        return 1 << self.layout_grid

    @property
    def side(self) -> float:
        return self.cells * self.cell_size

    @property
    def max_truncation(self) -> float:
        return self.truncation_cells * self.cell_size

    def world(self) -> World:
        c = self.side / 2.0
        return World.room_with_obstacles(c, c, self.room[0], self.room[1], self.n_obstacles, self.seed,
                                         obstacle_scale=self.obstacle_scale)

    def trajectory(self, n: int | None = None) -> np.ndarray:
        """(n, 3) ground-truth x, y, theta: start at the grid centre (ThreadLocalize.cpp:466-471), constant twist."""
        n = self.n_scans if n is None else n
        out = np.zeros((n, 3))
        x = y = self.side / 2.0
        th = 0.0
        for i in range(n):
            out[i] = (x, y, th)
            x += self.twist[0] * math.cos(th) - self.twist[1] * math.sin(th)
            y += self.twist[0] * math.sin(th) + self.twist[1] * math.cos(th)
            th += self.twist[2]
        return out

    def scans(self, n: int | None = None):
        """Yield (pose_xytheta, float32 ranges) along the trajectory."""
        world = self.world()
        rng = np.random.default_rng(self.seed + 7)
        for x, y, th in self.trajectory(n):
            yield (x, y, th), scan_from_pose(world, self.sensor, x, y, th, rng)


def config(which: str) -> Config:
    if which == "C1":
        return Config("C1", 10, seed=BASE_SEED + 1, room=(20.0, 15.0), n_obstacles=8)
    if which == "C2":
        return Config("C2", 12, seed=BASE_SEED + 2, room=(90.0, 67.5), n_obstacles=64,
                      sensor=SensorSpec(max_range=80.0))
    if which == "C3":
        # 1024 obstacles of the size of C1's (pillars, vehicles: 0.15 .. 0.6 m circumradius) in a 360 x 270 m hall: the
        # sensor sees most of the hall through them, every obstacle casts a shadow wedge, and most partitions in range
        # are "active" (addTsd per cell) rather than "empty" -- the addTsd-dominated regime of the large map.  (With
        # obstacles that grow with the room, 2 .. 8 m, the view ends after 30 m and a push touches 0.35 M cells.)
        return Config("C3", 14, seed=BASE_SEED + 3, room=(360.0, 270.0), n_obstacles=1024,
                      sensor=SensorSpec(max_range=250.0), n_scans=8, obstacle_scale=1.0)
    if which == "C5":
        return Config("C5", 16, seed=BASE_SEED + 5, room=(1500.0, 1200.0), n_obstacles=1024,
                      sensor=SensorSpec(max_range=800.0), n_scans=4)
    if which == "tiny":  # 256x256 cells, for second-scale CPU tests
        return Config("tiny", 8, seed=BASE_SEED, room=(5.0, 4.0), n_obstacles=3,
                      sensor=SensorSpec(beams=361, angular_res=math.pi / 240.0, max_range=8.0), n_scans=12)
    raise KeyError(which)
