"""Benchmark workload of bench.py: BASELINE.json configs[1] -- the double-laser configuration (two 1081-beam
scanners with local offsets +-0.35 m sharing one 4096 x 4096 grid, reference src/SlamNode.cpp:104-121), in the
DENSE regime (every partition already allocated, BASELINE.md section 2 "observation"): the map is first built
by pushing scans along the trajectory, then every partition that is still unallocated is allocated as free
space, so that each timed push rewrites every in-range visible cell (the bandwidth-bound regime the north
star's HBM target refers to).  Plain numpy; the same scans feed the CUDA arm and the reference arm."""
from __future__ import annotations

import math

import numpy as np

from . import synth
from .scan import HostSensor


class DoubleLaserWorkload:
    def __init__(self, config_name: str = "C2", n_map: int = 6, n_steps: int = 4, invert=None, seed_offset: int = 0,
                 origin=None):
        """origin: (x, y) of the robot's start and of its room's centre; default = the grid centre."""
        self.cfg = cfg = synth.config(config_name)
        self.name = config_name
        self.offsets = (+0.35, -0.35)
        rng = np.random.default_rng(cfg.seed + 7 + seed_offset)
        traj = cfg.trajectory(n_map + n_steps)
        if origin is None:
            self.world = cfg.world()
        else:
            self.world = synth.World.room_with_obstacles(origin[0], origin[1], cfg.room[0], cfg.room[1], cfg.n_obstacles,
                                                         cfg.seed + seed_offset)
            traj = traj + np.array([origin[0] - cfg.side / 2.0, origin[1] - cfg.side / 2.0, 0.0])
        self.sensors = [HostSensor(cfg.sensor, invert) for _ in self.offsets]
        self.map_scans = []    # list of Scan (alternating lasers)
        self.step_scans = []   # list of (ScanA, ScanB)
        self.step_rays = []
        for i, (x, y, th) in enumerate(traj):
            pair = []
            rays = []
            for hs, off in zip(self.sensors, self.offsets):
                sx = x - off * math.sin(th)
                sy = y + off * math.cos(th)
                r = synth.scan_from_pose(self.world, cfg.sensor, sx, sy, th, rng)
                hs.set_scan(r)
                hs.T = np.eye(3)
                hs.rays = hs.rays_local.copy()
                hs.ray_norm = 1.0
                hs.transform(synth.pose_matrix(sx, sy, th))
                pair.append(hs.scan())
                rays.append(hs.normalized_rays(cfg.cell_size).copy())
            if i < n_map:
                self.map_scans.extend(pair)
            else:
                self.step_scans.append(tuple(pair))
                self.step_rays.append(tuple(rays))

    def describe(self) -> dict:
        c = self.cfg
        return {
            "workload": f"{self.name}: double-laser (2 x {c.sensor.beams} beams, 270 deg, max range {c.sensor.max_range:g} m) on a "
                        f"{c.cells}x{c.cells} TsdGrid @ {c.cell_size * 100:g} cm, 32x32 partitions, dense regime "
                        f"(all {(c.cells // 32) ** 2} partitions allocated), room {c.room[0]:g}x{c.room[1]:g} m + {c.n_obstacles} obstacles",
            "grid_cells": c.cells * c.cells,
            "cell_state_bytes": (c.cells // 32) ** 2 * 8832 * 2,
            "pushes_per_step": 2,
            "l2_policy": f"cell state ({(c.cells // 32) ** 2 * 8832 * 2 / 1e6:.0f} MB) is larger than L2 (126 MB): inputs larger than L2, no flush",
        }

    def build_map(self, grid):
        """Same sequence on any backend with .push/.fill: build, then allocate the rest as free space."""
        for sc in self.map_scans:
            grid.push(sc)
        grid.fill(1.0, 1.0, only_uninitialized=True)


class MultiRobotWorkload:
    """bench.py at N > 1 GPUs (BASELINE.json configs[4], "multi-SLAM" on a grid sharded in bands): N double-laser
    robots of the C2 kind, each with its own room, on ONE TsdGrid of (8192 N)^2 cells sharded in N bands of 8192
    cell rows (256 partition rows; N = 8 is the 65536^2 grid of configs[4]).  Robot k starts 20 m above the lower
    edge of band k, so its room (67.5 m tall) straddles the boundary to band k-1: the lower 14 m of every room
    are integrated by the neighbouring GPU and the boundary rows really change every step.  Per-GPU work is one
    robot's worth for every N: weak scaling."""

    BAND_ROWS = 8192
    RISE = 20.0  # metres between a band's lower edge and its robot's start

    def __init__(self, n_robots: int, config_name: str = "C2", n_map: int = 6, n_steps: int = 4, invert=None):
        base = synth.config(config_name)
        self.n = n_robots
        cells = self.BAND_ROWS * n_robots
        self.layout_grid = int(round(math.log2(cells)))
        if (1 << self.layout_grid) != cells:
            raise ValueError("the number of GPUs must be a power of two (the reference's grids are 2^k cells wide)")
        self.cell_size = base.cell_size
        self.max_truncation = base.max_truncation
        side = cells * base.cell_size
        band_h = self.BAND_ROWS * base.cell_size
        self.robots = [DoubleLaserWorkload(config_name, n_map, n_steps, invert, seed_offset=1000 * k,
                                           origin=(side / 2.0, k * band_h + self.RISE)) for k in range(n_robots)]
        self.cfg = base
        self.sensors = self.robots[0].sensors
        # one step = every robot's two scans, robot after robot
        self.map_scans = [sc for i in range(0, 2 * n_map, 2) for r in self.robots for sc in r.map_scans[i:i + 2]]
        self.step_scans = [tuple(sc for r in self.robots for sc in r.step_scans[i]) for i in range(n_steps)]
        self.step_rays = [tuple(ry for r in self.robots for ry in r.step_rays[i]) for i in range(n_steps)]

    def describe(self) -> dict:
        c = self.cfg
        cells = 1 << self.layout_grid
        band_bytes = (self.BAND_ROWS // 32 + 2) * (cells // 32) * 8832 * 2
        return {
            "workload": f"{self.n} x C2 robot (double-laser, 2 x {c.sensor.beams} beams, max range {c.sensor.max_range:g} m, room "
                        f"{c.room[0]:g}x{c.room[1]:g} m + {c.n_obstacles} obstacles each) on one {cells}x{cells} TsdGrid @ "
                        f"{c.cell_size * 100:g} cm sharded in {self.n} bands of {self.BAND_ROWS} cell rows, dense regime; every room "
                        f"straddles a band boundary",
            "grid_cells": cells * cells,
            "cell_state_bytes_per_gpu": band_bytes,
            "pushes_per_step": 2 * self.n,
            "l2_policy": f"cell state per GPU ({band_bytes / 1e9:.1f} GB) is larger than L2 (126 MB): inputs larger than L2, no flush",
        }
