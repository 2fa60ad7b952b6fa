"""Host-side pieces: synthetic data generator determinism, the host sensor mirror, gslcblas-order gemm."""
import os

import numpy as np

from ohm_tsd_slam_b200 import synth
from ohm_tsd_slam_b200.scan import HostSensor, gemm_nn, standard_mask
from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_is_deterministic_and_hokuyo_like():
    a = list(synth.config("C1").scans(3))
    b = list(synth.config("C1").scans(3))
    for (pa, ra), (pb, rb) in zip(a, b):
        assert pa == pb and np.array_equal(ra, rb)
    r = a[0][1]
    assert r.dtype == np.float32 and len(r) == 1081
    assert np.isinf(r).sum() > 0 and (r == 0).sum() > 0
    fin = r[np.isfinite(r) & (r > 0)]
    assert fin.min() > 0.1 and fin.max() <= 30.0


def test_gemm_nn_skips_zero_coefficients():
    A = np.array([[0.0, 2.0], [3.0, 0.0]])
    B = np.array([[np.inf, 1.0], [1.0, np.nan]])
    C = gemm_nn(A, B)
    # 0 * inf and 0 * nan are never formed
    assert C[0, 0] == 2.0 and np.isnan(C[0, 1]) and np.isinf(C[1, 0]) and C[1, 1] == 3.0


def test_standard_mask_rules():
    spec = synth.SensorSpec(beams=9, angular_res=np.pi / 180, max_range=10.0)
    data = np.array([1.0, 0.0, np.nan, 20.0, 1.0, 1.0, 5.0, 1.0, 1.0])
    d, m = standard_mask(data, spec)
    assert m[1] == 0 and m[2] == 0 and np.isinf(d[2]) and np.isinf(d[3]) and m[3] == 1
    assert m[6] == 0  # 5 m between 1 m neighbours: acute angle -> depth discontinuity


def test_host_sensor_scene_and_rays():
    cfg = synth.config("tiny")
    hs = HostSensor(cfg.sensor, port.invert3x3)
    (x, y, th), r = next(iter(cfg.scans(1)))
    hs.set_scan(r)
    hs.transform(synth.pose_matrix(x, y, th))
    rays = hs.normalized_rays(cfg.cell_size)
    assert np.allclose(np.hypot(rays[0], rays[1]), cfg.cell_size, rtol=1e-12)
    coords, mask, n = hs.scene()
    assert n == int(mask.sum()) and np.allclose(np.hypot(*coords[mask > 0].T), hs.data[mask > 0], rtol=1e-12)
    sc = hs.scan()
    assert np.allclose(sc.pose_inv @ sc.pose, np.eye(3), atol=1e-12)


def test_build_stamp_follows_the_sources_not_their_mtimes(monkeypatch):
    """The library is rebuilt when the content hash of its sources differs from the stamp written at build time
    (a snapshot copied to another box does not keep the order of modification times)."""
    from ohm_tsd_slam_b200 import _build, capi
    capi.lib()  # built and stamped
    assert not _build.needs_build()
    import os
    os.utime(_build.sources()[0])  # a newer mtime alone changes nothing
    assert not _build.needs_build()
    monkeypatch.setattr(_build, "source_hash", lambda: "something else")
    assert _build.needs_build()


def test_fast_path_front_end_is_exact_where_it_is_certain(tmp_path):
    """k_update's single-precision front end (csrc/beam_index.cuh: tsd_fast_model / tsd_gate_entry / tsd_classify_cell,
    host/device code) against the reference's double-precision expressions (SensorPolar2D.cpp:117-135, TsdGrid.cpp:250-274,
    TsdGridPartition.h:174-191) on the CPU: every cell the front end decides must be decided as the reference decides
    it -- random and adversarial cells (next to beam boundaries, next to the truncation band, on the cut of atan2),
    approximate reciprocal perturbed by up to +-2 ulp.  tests/cpp/fastpath_check.cpp."""
    import subprocess
    exe = str(tmp_path / "fastpath_check")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", os.path.join(ROOT, "tests", "cpp", "fastpath_check.cpp"),
                    "-o", exe], check=True)
    out = subprocess.run([exe, "1500000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "violations 0" in out.stdout
    # the front end must decide almost everything, or it is not a fast path
    certain = float(out.stdout.split("certain ")[1].split("%")[0])
    assert certain > 99.0, out.stdout


def test_raycast_closed_form_positions_equal_the_serial_sums(tmp_path):
    """k_raycast replaces the 32 serial additions `position += ray` of a pass by p + k d where that is exact
    (ohm_tsd_slam_b200/csrc/closed_form.cuh, host/device code).  tests/cpp/closedform_check.cpp runs that very function on
    the CPU against the serial sums over 12 M cases (all binades of the map, binade boundaries, exact multiples and exact
    ties of the position's ulp, tiny and zero steps): whenever it says "closed", all 32 partial sums agree bit for bit."""
    import subprocess
    exe = str(tmp_path / "closedform_check")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Werror",
                    os.path.join(ROOT, "tests", "cpp", "closedform_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 mismatches" in out.stdout and "400000 exact ties refused" in out.stdout
    closed = float(out.stdout.split("(")[1].split("%")[0])
    assert closed > 50.0, out.stdout  # (an adversarial mix: a third of the cases sit on binade boundaries, ties or zero)


def test_rnm_box_bound_never_exceeds_a_true_distance(tmp_path):
    """k_score_rnm skips a group of model points when a single-precision lower bound of the squared distance to the
    group's bounding box exceeds the best distance so far (ohm_tsd_slam_b200/csrc/nn_bounds.cuh, host/device code).
    tests/cpp/nnbound_check.cpp runs those very functions on the CPU: over 2.4 M queries (anywhere, just outside an edge,
    inside, off a corner; coordinates up to the largest map) the bound never exceeds the double-precision squared
    distance to any point of the box, rounded up to single precision -- the quantity the kernel compares it with."""
    import subprocess
    exe = str(tmp_path / "nnbound_check")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Werror",
                    os.path.join(ROOT, "tests", "cpp", "nnbound_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 violations" in out.stdout
    assert float(out.stdout.split("positive for ")[1].split("%")[0]) > 40.0  # (the bound is not trivially zero)


def test_icp_cell_gaps_never_exceed_a_true_distance(tmp_path):
    """k_icp opens a neighbouring hash cell only if the query's gap to it does not exceed the best squared distance so far
    (ohm_tsd_slam_b200/csrc/icp_cells.cuh, host/device code).  tests/cpp/icpcell_check.cpp runs those very functions on
    the CPU: over 15 M query / model point pairs (cell boundaries included, origins kilometres away, five cell sizes) a
    point lying in a neighbouring cell is never nearer than the gap says."""
    import subprocess
    exe = str(tmp_path / "icpcell_check")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Werror",
                    os.path.join(ROOT, "tests", "cpp", "icpcell_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 violations" in out.stdout
