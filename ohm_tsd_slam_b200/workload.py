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
    def __init__(self, config_name: str = "C2", n_map: int = 6, n_steps: int = 4, invert=None, seed_offset: int = 0):
        self.cfg = cfg = synth.config(config_name)
        self.name = config_name
        self.world = cfg.world()
        self.offsets = (+0.35, -0.35)
        rng = np.random.default_rng(cfg.seed + 7 + seed_offset)
        traj = cfg.trajectory(n_map + n_steps)
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
