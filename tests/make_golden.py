"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libohm_ref.so = the reference's own
sources compiled unmodified, see oracle/Makefile).  Run where /root/reference exists:

    make -C oracle ref && python tests/make_golden.py

The fixtures pin the port oracle (tests/test_oracle_pin.py) and, through it, the CUDA path on boxes that
have neither /root/reference nor the prebuilt reference library.
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ohm_tsd_slam_b200 import synth
from oracle import ref

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_sequence(name: str, n_scans: int, parts_to_dump: int):
    cfg = synth.config(name)
    ref.set_threads(1)
    g = ref.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
    g.set_max_truncation(cfg.max_truncation)
    s = ref.Sensor(cfg.sensor)
    icp = ref.Icp(30, 0.4, 0.02, g.bounds)
    scans = list(cfg.scans(n_scans))
    (x, y, th), r0 = scans[0]
    s.set_scan(r0)
    s.transform(synth.pose_matrix(x, y, th))
    assert g.free_footprint(x, y, 0.6, 0.6)
    g.push(s)
    out = {"n_scans": n_scans}
    rng = np.random.default_rng(99)
    for k, (_, r) in enumerate(scans[1:]):
        s.set_scan(r)
        out[f"data_{k}"] = s.data
        out[f"mask_{k}"] = s.mask
        out[f"pose_{k}"] = s.pose
        out[f"pose_inv_{k}"] = ref.invert(s.pose)
        coords, normals, mask, cnt = g.raycast_mask(s)
        out[f"rays_{k}"] = s.normalized_rays(cfg.cell_size)
        out[f"rc_coords_{k}"] = np.where(mask[:, None] > 0, coords, 0.0)
        out[f"rc_normals_{k}"] = np.where(mask[:, None] > 0, normals, 0.0)
        out[f"rc_mask_{k}"] = mask
        scene, mask_s, _ = s.scene()
        Mv, Nv, Sv = coords[mask > 0], normals[mask > 0], scene[mask_s > 0]
        its, pm, ps, pc, rms, Tf = icp.trace(Mv, Nv, Sv, s.pose)
        T, rmsv, pairs, it, st = icp.run(Mv, Nv, Sv, s.pose)
        out[f"icp_T_{k}"] = T
        out[f"icp_stats_{k}"] = np.array([rmsv, pairs, it, st])
        out[f"icp_pair_count_{k}"] = pc[:its]
        out[f"icp_pairs_model_{k}"] = np.concatenate([pm[i, :pc[i]] for i in range(its)]).astype(np.uint16)
        out[f"icp_pairs_scene_{k}"] = np.concatenate([ps[i, :pc[i]] for i in range(its)]).astype(np.uint16)
        assert np.array_equal(Tf[its - 1][:2, :2], T[:2, :2])
        s.transform(T)
        g.push(s)
        out[f"pose_after_{k}"] = s.pose
        st_, iw = g.partition_states()
        out[f"states_{k}"] = st_.astype(np.uint8)
        out[f"initw_{k}"] = iw
    # final cell state of a deterministic sample of partitions + a checksum over all of them
    st_, iw = g.partition_states()
    content = np.nonzero(st_ == 2)[0]
    pick = content[np.linspace(0, len(content) - 1, min(parts_to_dump, len(content))).astype(int)]
    out["dump_parts"] = pick.astype(np.int32)
    tsd = np.stack([g.download_partition(int(p))[0] for p in pick])
    wgt = np.stack([g.download_partition(int(p))[1] for p in pick])
    out["dump_tsd"] = tsd
    out["dump_weight"] = wgt
    out["checksum"] = grid_checksum(g, st_)
    xy = rng.uniform(-0.2, cfg.side + 0.2, size=(4000, 2))
    t, stt = g.interpolate_bilinear(xy)
    out["interp_xy"], out["interp_tsd"], out["interp_status"] = xy, t, stt.astype(np.uint8)
    nn, ok = g.interpolate_normal(xy)
    out["normal_n"], out["normal_ok"] = np.where(ok[:, None] > 0, nn, 0.0), ok.astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, f"sequence_{name}.npz"), **out)
    print(name, "written", {k: v for k, v in zip(*np.unique(st_, return_counts=True))})
    return cfg, g, s, scans


def grid_checksum(g, states) -> np.ndarray:
    """Order-independent 64-bit checksum (sum of the raw bit patterns mod 2^64) of all cells of all content partitions."""
    acc_t = np.uint64(0)
    acc_w = np.uint64(0)
    with np.errstate(over="ignore"):
        for p in np.nonzero(states == 2)[0]:
            t, w = g.download_partition(int(p))
            t = np.where(np.isnan(t), np.float64("nan"), t)  # canonical NaN
            acc_t += t.view(np.uint64).sum(dtype=np.uint64) * np.uint64(int(p) * 2 + 1)
            acc_w += w.view(np.uint64).sum(dtype=np.uint64) * np.uint64(int(p) * 2 + 1)
    return np.array([acc_t, acc_w], dtype=np.uint64)


def run_matchers(name: str):
    cfg = synth.config(name)
    ref.set_threads(1)
    g = ref.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
    g.set_max_truncation(cfg.max_truncation)
    s = ref.Sensor(cfg.sensor)
    scans = list(cfg.scans(4))
    (x, y, th), r0 = scans[0]
    s.set_scan(r0)
    s.transform(synth.pose_matrix(x, y, th))
    g.push(s)
    icp = ref.Icp(30, 0.4, 0.02, g.bounds)
    out = {}
    res = cfg.sensor.angular_res
    phimax = math.radians(30.0)
    for k, (_, r) in enumerate(scans[1:]):
        s.set_scan(r)
        coords, normals, mask, cnt = g.raycast_mask(s)
        scene, mask_s, _ = s.scene()
        coords = np.where(mask[:, None] > 0, coords, 0.0)
        out[f"M_{k}"], out[f"maskM_{k}"], out[f"S_{k}"], out[f"maskS_{k}"] = coords, mask, scene, mask_s
        out[f"pose_{k}"] = s.pose
        for trials, ctrl in ((20, 60), (100, 140)):
            tag = f"{k}_{trials}_{ctrl}"
            ref.seed(1000 + k)
            out[f"tsd_{tag}"] = ref.match_tsd(g, trials, 0.15, ctrl, 0.25, s.pose, coords, mask, scene, mask_s, phimax, 0.25, res)
            ref.seed(2000 + k)
            out[f"rnm_{tag}"] = ref.match_rnm(trials, 0.15, ctrl, coords, mask, scene, mask_s, phimax, 0.25, res)
            ref.seed(3000 + k)
            out[f"pdf_{tag}"] = ref.match_pdf(trials, 0.15, ctrl, ref.PDF_DEFAULTS, coords, mask, scene, mask_s, phimax, 0.25, res)
        T = icp.run(coords[mask > 0], normals[mask > 0], scene[mask_s > 0], s.pose)[0]
        s.transform(T)
        g.push(s)
    # the map the TSD matcher looked at is reproduced in the test by replaying the same pushes
    np.savez_compressed(os.path.join(OUT, f"matchers_{name}.npz"), **out)
    print(name, "matchers written")


def run_axis_map(name: str, n_scans: int):
    """Map publication (ThreadGrid.cpp:84,125): RayCastAxisAligned2D::calcCoords and TsdGrid::grid2ColorImage of the
    reference after n_scans pushes at the ground-truth poses (tests/harness.py axis_map_scenario replays them)."""
    from tests.harness import axis_map_scenario
    cfg = synth.config(name)
    ref.set_threads(1)
    g = ref.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
    g.set_max_truncation(cfg.max_truncation)
    s = ref.Sensor(cfg.sensor)
    for sc in axis_map_scenario(cfg, n_scans, ref.invert):
        s.set_data(sc.ranges, sc.mask)
        s.pose = sc.pose
        g.push(s)
    coords, normals, occ = g.axis_map(with_normals=True)
    img = g.color_image(320, 200)
    np.savez_compressed(os.path.join(OUT, f"axis_map_{name}.npz"), n_scans=n_scans, coords=coords, normals=normals,
                        occupied=occ, image=img)
    print(name, "axis map written:", len(coords), "crossings,", int((occ == 0).sum()), "free cells")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "axis":  # one reference grid per process (see oracle/README.md)
        run_axis_map(sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    run_sequence("tiny", 10, 12)
    run_sequence("C1", 6, 6)
    run_matchers("tiny")
