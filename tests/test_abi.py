"""The C-ABI shared library loads and exports every symbol include/tsdslam_b200.h declares; without a CUDA
device the compute entry points fail loudly (TSD_E_NO_DEVICE) -- there is no CPU path.  CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ohm_tsd_slam_b200 import capi
from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "tsdslam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n.startswith(("tsd", "icp_", "match_"))))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = declared_functions()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(set(capi.EXPORTS)) == names


def test_invert3x3_matches_oracle():
    rng = np.random.default_rng(1)
    for _ in range(50):
        th = rng.uniform(-3, 3)
        T = np.array([[np.cos(th), -np.sin(th), rng.uniform(0, 100)], [np.sin(th), np.cos(th), rng.uniform(0, 100)], [0, 0, 1.0]])
        assert np.array_equal(capi.invert3x3(T), port.invert3x3(T))
        assert np.allclose(capi.invert3x3(T) @ T, np.eye(3), atol=1e-12)


def test_no_cpu_fallback():
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = capi.lib().tsdg_create(0.025, 5, 8, 0, C.byref(h))
    assert rc == -2 and not h.value  # TSD_E_NO_DEVICE
    assert b"no CPU path" in capi.lib().tsd_last_error()
    with pytest.raises(capi.TsdError):
        capi.Grid(0.025, 5, 8)
    with pytest.raises(capi.TsdError):
        capi.Icp(30, 0.4, 0.02, (0, 1, 0, 1))
    with pytest.raises(capi.TsdError):
        capi.Matcher()


def test_argument_validation():
    h = C.c_void_p()
    assert capi.lib().tsdg_create(0.025, 4, 8, 0, C.byref(h)) in (-1, -2)  # only 32x32 partitions
    assert capi.lib().tsdg_create(0.025, 5, 20, 0, C.byref(h)) in (-1, -2)
    assert capi.lib().tsdg_push(None, None) == -1
    assert capi.lib().tsdg_destroy(None) == 0


def test_argument_validation_of_the_wider_surface(tmp_path):
    """NULL handles / buffers are refused before anything touches CUDA (runs without a GPU)."""
    L = capi.lib()
    n = C.c_uint32()
    assert L.tsdg_push_batch(None, None, 2) == -1
    assert L.tsdg_push_batch_async(None, None, 0) == -1
    assert L.tsdg_stage_batch(None, None, 1) == -1
    assert L.tsdg_axis_aligned_map(None, None, 0, None, C.byref(n), None) == -1
    assert L.tsdg_color_image(None, None, 4, 4) == -1
    assert L.tsdg_store(None, b"x") == -1
    h = C.c_void_p()
    assert L.tsdg_load(None, 0, C.byref(h)) == -1
    assert L.tsdg_load(str(tmp_path / "missing.txt").encode(), 0, C.byref(h)) == -1 and not h.value
    assert b"cannot open" in L.tsd_last_error()
    assert L.tsdg_band_halo_sync(None, 0, -1, 0, -1) == -1
    assert L.tsdg_band_push_finish(None) == -1
    assert L.tsdg_band_export(None, None) == -1
    assert L.tsdg_band_connect(None, 0, None) == -1
    box = (C.c_int32 * 4)()
    assert L.tsdg_scan_box(None, None, box) == -1


def test_argument_validation_of_the_round_two_surface():
    """The entry points added in round 2 refuse NULL handles / buffers and bad sizes before anything touches CUDA."""
    L = capi.lib()
    h = C.c_void_p()
    one = (C.c_int * 1)(0)
    assert L.tsdg_create_sharded(0.025, 5, 8, 0, one, C.byref(h)) == -1 and not h.value      # no bands
    assert L.tsdg_create_sharded(0.025, 5, 8, 99, one, C.byref(h)) == -1 and not h.value     # more bands than partition rows
    assert L.tsdg_create_sharded(0.025, 5, 3, 1, one, C.byref(h)) == -1                      # invalid layout
    assert L.tsdg_create_sharded(0.025, 5, 8, 1, one, None) == -1
    assert L.tsdg_sharded_destroy(None) == 0
    assert L.tsdg_sharded_num_bands(None) == 0
    assert not L.tsdg_sharded_band(None, 0)
    assert L.tsdg_sharded_push(None, None) == -1
    assert L.tsdg_sharded_push_batch(None, None, 2) == -1
    assert L.tsdg_sharded_sync(None) == -1
    assert L.tsdg_sharded_set_max_truncation(None, 0.1) == -1
    assert L.tsdg_sharded_free_footprint(None, 0.0, 0.0, 1.0, 1.0) == -1
    assert L.tsdg_sharded_raycast_mask(None, None, None, None, None, None, None) == -1
    assert L.tsdg_sharded_interpolate_bilinear(None, 1, None, None, None) == -1
    assert L.tsdg_sharded_partition_states(None, None, None) == -1
    assert L.tsdg_sharded_download_partition(None, 0, None, None) == -1
    assert L.tsdg_localize(None, None, None, None, None, 0, None, None, None, None, None, None, None) == -1
    assert L.icp_pairs(None, None, 0, None, 0, None, None, None, None, None) == -1
    assert L.match_prepare(None, 10, None, None, None, None, 10, 360, 30, 0.5, 0.004, 1, None) == -1
    assert L.tsds_prepare_scan(None, 0, None, 1.0, 30.0, 0.004, None, None, None, None, None, None, None) == -1
    assert L.tsdg_raycast_mask_sharded(None, None, None, None, None, None, None) == -1
    assert L.tsdg_raycast_sharded_launch(None, None, None) == -1
    assert L.tsdg_raycast_sharded_collect(None, 0, None, None, None, None) == -1
    assert L.tsdg_band_rcx_export(None, None) == -1
    # the counter-based generator of match_prepare is a pure host/device function: fixed values, uniform-looking residues
    assert capi.match_rng(1, 2, 3) == 0x12e80ffcfbbbd251
    r = [capi.match_rng(7, 0, i) % 1000 for i in range(4000)]
    assert 450 < sum(r) / len(r) < 550 and len(set(r)) > 900
