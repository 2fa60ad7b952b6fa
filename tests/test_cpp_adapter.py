"""The header-compatible C++ mirror of the obvious:: interface (ohm_tsd_slam_b200/obvious/): tests/cpp/slam_loop.cpp
replays ThreadLocalize / ThreadMapping against it.  The same source compiled against the reference's own headers
produced tests/golden/slam_loop_*_mode0.txt (see the header of slam_loop.cpp)."""
import math
import os
import subprocess

import numpy as np
import pytest

from ohm_tsd_slam_b200 import _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "slam_loop_b200")
PHI_MIN = repr(-135.0 * math.pi / 180.0)
ARGS = {
    "tiny": ["8", "0.025", "3", "361", repr(math.pi / 240.0), PHI_MIN, "8.0", "0.001", "2.0"],
    "C1": ["10", "0.025", "3", "1081", repr(math.pi / 720.0), PHI_MIN, "30.0", "0.001", "2.0"],
}


def build_exe():
    _build.build()
    src = os.path.join(ROOT, "tests", "cpp", "slam_loop.cpp")
    hdr = os.path.join(ROOT, "ohm_tsd_slam_b200", "obvious", "obvious_b200.h")
    if os.path.exists(EXE) and os.path.getmtime(EXE) > max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(_build.LIB)):
        return EXE
    libdir = os.path.join(ROOT, "ohm_tsd_slam_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I" + os.path.join(libdir, "obvious"), src, "-o", EXE,
                    "-L" + libdir, "-ltsdslam_b200", "-Wl,-rpath," + libdir], check=True)
    return EXE


def run(name, mode, bands=None, device_prep=False):
    env = dict(os.environ)
    if bands:
        env["SLAM_LOOP_BANDS"] = str(bands)
    if device_prep:
        env["SLAM_LOOP_DEVICE_PREP"] = "1"
    out = subprocess.run([build_exe(), os.path.join(ROOT, "tests", "golden", f"scans_{name}.bin")] + ARGS[name] + [str(mode)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr
    return np.array([[float(v) for v in line.split()] for line in out.stdout.strip().splitlines()])


def test_node_code_compiles_against_adapter_headers():
    assert os.path.exists(build_exe())


def test_adapter_fails_loudly_without_gpu():
    from ohm_tsd_slam_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    out = subprocess.run([build_exe(), os.path.join(ROOT, "tests", "golden", "scans_tiny.bin")] + ARGS["tiny"] + ["0"],
                         capture_output=True, text=True)
    assert out.returncode != 0 and "no CPU path" in out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny", "C1"])
def test_slam_loop_tracks_reference_poses(name):
    got = run(name, 0)
    ref = np.loadtxt(os.path.join(ROOT, "tests", "golden", f"slam_loop_{name}_mode0.txt"))
    assert got.shape == ref.shape
    # the last line is the map publication summary (ThreadGrid), the others one pose per scan
    pub_got, pub_ref, got, ref = got[-1], ref[-1], got[:-1], ref[:-1]
    assert np.array_equal(got[:, [0, 4, 6]], ref[:, [0, 4, 6]])       # scan index, model points, iterations
    assert np.max(np.abs(got[:, 1:4] - ref[:, 1:4])) < 1e-8             # x, y, theta after every scan
    assert np.max(np.abs(got[:, 5] - ref[:, 5])) <= 2                   # ICP pairs (ulp-level pose differences)
    # RayCastAxisAligned2D::calcCoords + grid2ColorImage on a map that differs from the reference's by those ulps:
    # crossings, their coordinate sums, free cells of the occupancy grid, byte sum of the colour image
    assert pub_got[0] == -1 and abs(pub_got[1] - pub_ref[1]) <= 2
    assert np.max(np.abs(pub_got[2:4] - pub_ref[2:4]) / pub_ref[2:4]) < 1e-3
    assert abs(pub_got[4] - pub_ref[4]) <= 8 and abs(pub_got[5] - pub_ref[5]) / pub_ref[5] < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name,bands", [("tiny", 2), ("C1", 4)])
def test_slam_loop_on_a_sharded_grid_is_identical(name, bands):
    """obvious::TsdGrid(cellSize, layoutPartition, layoutGrid, nBands, devices): the node's loop on a grid sharded inside
    the library prints the same poses, digit for digit, as on the unsharded grid."""
    a = run(name, 0)[:-1]
    b = run(name, 0, bands=bands)
    assert a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.gpu
def test_slam_loop_with_tsd_matcher():
    """registration_mode 3 (TSD_PDFMatching pre-registration, random control set): poses stay with the ICP-only run."""
    a = run("tiny", 0)[:-1]
    b = run("tiny", 3)[:-1]
    assert a.shape == b.shape and np.max(np.abs(a[:, 1:3] - b[:, 1:3])) < 0.05 and np.max(np.abs(a[:, 3] - b[:, 3])) < 0.02
    # the same with the matcher's pre-processing on the device (match_prepare, counter-based random numbers)
    c = run("tiny", 3, device_prep=True)[:-1]
    assert a.shape == c.shape and np.max(np.abs(a[:, 1:3] - c[:, 1:3])) < 0.05 and np.max(np.abs(a[:, 3] - c[:, 3])) < 0.02


THREADS_EXE = os.path.join(ROOT, "tests", "cpp", "threads_b200")


def build_threads_exe():
    _build.build()
    src = os.path.join(ROOT, "tests", "cpp", "threads.cpp")
    if os.path.exists(THREADS_EXE) and os.path.getmtime(THREADS_EXE) > max(os.path.getmtime(src), os.path.getmtime(_build.LIB)):
        return THREADS_EXE
    libdir = os.path.join(ROOT, "ohm_tsd_slam_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-pthread", src, "-o", THREADS_EXE, "-L" + libdir,
                    "-ltsdslam_b200", "-Wl,-rpath," + libdir], check=True)
    return THREADS_EXE


def test_thread_test_compiles():
    assert os.path.exists(build_threads_exe())


@pytest.mark.gpu
def test_one_handle_from_three_threads_equals_the_serial_run():
    """The node's threading (1 mapper + 2 localisers on one TsdGrid, no lock in the reference: ThreadMapping.cpp:46-61,
    SlamNode.cpp:104-121) against the C ABI: 500 iterations of push / push_async on one thread and ray casts + samples
    on two more, results and final map bit-identical to the single-threaded run (tests/cpp/threads.cpp)."""
    out = subprocess.run([build_threads_exe(), "500"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "identical to the single-threaded run" in out.stdout


MATCH_EXE = os.path.join(ROOT, "tests", "cpp", "matchers_adapter_b200")


def build_matchers_exe():
    _build.build()
    src = os.path.join(ROOT, "tests", "cpp", "matchers_adapter.cpp")
    hdr = os.path.join(ROOT, "ohm_tsd_slam_b200", "obvious", "obvious_b200.h")
    if os.path.exists(MATCH_EXE) and os.path.getmtime(MATCH_EXE) > max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(_build.LIB)):
        return MATCH_EXE
    libdir = os.path.join(ROOT, "ohm_tsd_slam_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Werror", "-I" + os.path.join(libdir, "obvious"), src, "-o",
                    MATCH_EXE, "-L" + libdir, "-ltsdslam_b200", "-Wl,-rpath," + libdir], check=True)
    return MATCH_EXE


def test_matcher_replay_compiles():
    assert os.path.exists(build_matchers_exe())


@pytest.mark.gpu
def test_adapter_matchers_reproduce_reference_golden(tmp_path):
    """obvious::TSD_PDFMatching / RandomNormalMatching / PDFMatching::match of the adapter (host pre-processing restated
    from RandomMatching.cpp:52-183 incl. Matrix::pcaAnalysis' Jacobi SVD, hypothesis scoring on the device) under the
    replayed rand() stream against the reference's own results (tests/golden/matchers_tiny.npz): the same 3x3
    transformation for every scan, matcher and parameter set.  Tolerance 1e-12 (the winner is the same hypothesis; its
    matrix is formed on the host from the same two angles)."""
    import struct
    from ohm_tsd_slam_b200 import synth
    G = np.load(os.path.join(ROOT, "tests", "golden", "matchers_tiny.npz"))
    cfg = synth.config("tiny")
    sp = cfg.sensor
    scans = list(cfg.scans(4))
    path = str(tmp_path / "matchers.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<3i", sp.beams, 3, cfg.layout_grid))
        f.write(struct.pack("<8d", cfg.cell_size, cfg.truncation_cells, sp.angular_res, sp.phi_min, sp.max_range, sp.min_range,
                            sp.low_reflectivity_range, math.radians(30.0)))
        (x, y, th), r0 = scans[0]
        f.write(np.asarray(r0, dtype=np.float32).tobytes())
        f.write(synth.pose_matrix(x, y, th).astype(np.float64).tobytes())
        for k in range(3):
            f.write(np.asarray(scans[k + 1][1], dtype=np.float32).tobytes())
            f.write(G[f"pose_{k}"].astype(np.float64).tobytes())
            f.write(np.ascontiguousarray(G[f"M_{k}"], dtype=np.float64).tobytes())
            f.write(np.ascontiguousarray(G[f"S_{k}"], dtype=np.float64).tobytes())
            f.write(G[f"maskM_{k}"].astype(np.uint8).tobytes())
            f.write(G[f"maskS_{k}"].astype(np.uint8).tobytes())
            after = G[f"pose_{k + 1}"] if k < 2 else np.zeros((3, 3))
            f.write(after.astype(np.float64).tobytes())
    out = subprocess.run([build_matchers_exe(), path], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert len(lines) == 18
    worst = 0.0
    exact = 0
    for ln in lines:
        p = ln.split()
        T = np.array([float(v) for v in p[4:]]).reshape(3, 3)
        ref_T = G[f"{p[0]}_{p[1]}_{p[2]}_{p[3]}"]
        assert not np.array_equal(ref_T, np.eye(3)) or p[0] != "tsd"
        worst = max(worst, float(np.max(np.abs(T - ref_T))))
        exact += int(np.array_equal(T, ref_T))
    assert worst < 1e-12, (worst, exact, out.stdout)
    assert exact >= 12, (exact, out.stdout)
