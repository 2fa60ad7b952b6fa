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


def run(name, mode):
    out = subprocess.run([build_exe(), os.path.join(ROOT, "tests", "golden", f"scans_{name}.bin")] + ARGS[name] + [str(mode)],
                         capture_output=True, text=True, timeout=300)
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
def test_slam_loop_with_tsd_matcher():
    """registration_mode 3 (TSD_PDFMatching pre-registration, random control set): poses stay with the ICP-only run."""
    a = run("tiny", 0)[:-1]
    b = run("tiny", 3)[:-1]
    assert a.shape == b.shape and np.max(np.abs(a[:, 1:3] - b[:, 1:3])) < 0.05 and np.max(np.abs(a[:, 3] - b[:, 3])) < 0.02


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
