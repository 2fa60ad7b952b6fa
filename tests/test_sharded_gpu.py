"""A grid sharded in bands must equal the unsharded grid bit for bit (SURVEY.md 7 hard part 7): several bands
on one device, boundary rows exchanged by device copies, ray crossings merged by min over the bands."""
import numpy as np
import pytest

from ohm_tsd_slam_b200 import capi, synth
from ohm_tsd_slam_b200.scan import HostSensor
from ohm_tsd_slam_b200.sharded import LocalBands
from tests.harness import compare_grids, same

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,bands,peer", [("tiny", 2, False), ("tiny", 3, True), ("C1", 2, True), ("C1", 5, False), ("C1", 4, True)])
def test_banded_grid_equals_single_grid(name, bands, peer):
    """peer=False: halo rows by device copies + tsdg_band_push_finish; peer=True: the one-kernel peer-memory
    synchronisation (tsdg_band_halo_sync), every band's kernel running concurrently on its own stream."""
    cfg = synth.config(name)
    whole = capi.Grid(cfg.cell_size, 5, cfg.layout_grid)
    parts = LocalBands(cfg.cell_size, cfg.layout_grid, bands, peer=peer)
    whole.set_max_truncation(cfg.max_truncation)
    parts.set_max_truncation(cfg.max_truncation)
    hs = HostSensor(cfg.sensor, capi.invert3x3)
    scans = list(cfg.scans(7))
    (x, y, th), r0 = scans[0]
    # a trajectory that crosses a band boundary: start next to it
    hs.set_scan(r0)
    hs.transform(synth.pose_matrix(x, y, th))
    assert whole.free_footprint(x, y, 0.6, 0.6) and parts.free_footprint(x, y, 0.6, 0.6)
    for k, (pose, r) in enumerate(scans):
        hs.set_scan(r)
        hs.T = synth.pose_matrix(*pose)
        sc = hs.scan()
        if k > 0:
            rays = hs.normalized_rays(cfg.cell_size).copy()
            c1, n1, m1, k1 = whole.raycast_mask(sc, rays)
            c2, n2, m2, k2 = parts.raycast_mask(sc, rays)
            assert same(m1, m2), (k, int((m1 != m2).sum()))
            assert same(c1[m1 > 0], c2[m2 > 0]) and same(n1[m1 > 0], n2[m2 > 0])
            assert k1 == k2 and k1 > 0
        whole.push(sc)
        parts.push(sc)
        a, b = whole.last_push_stats(), parts.last_push_stats()
        assert a["cell_updates"] == b["cell_updates"] and a["active_tiles"] == b["active_tiles"]
        ok, lines = compare_grids(whole, parts)
        assert ok, (k, lines)


def test_pushes_need_no_communication():
    """Several pushes from two sensor poses, ONE halo synchronisation at the end: same grid as unsharded."""
    cfg = synth.config("C1")
    whole = capi.Grid(cfg.cell_size, 5, cfg.layout_grid)
    parts = LocalBands(cfg.cell_size, cfg.layout_grid, 4, peer=True)
    whole.set_max_truncation(cfg.max_truncation)
    parts.set_max_truncation(cfg.max_truncation)
    hs = HostSensor(cfg.sensor, capi.invert3x3)
    scans = list(cfg.scans(6))
    for k, (pose, r) in enumerate(scans):
        hs.set_scan(r)
        hs.T = synth.pose_matrix(*pose)
        sc = hs.scan()
        whole.push(sc)
        parts.push(sc, sync=False)
    parts.sync_halos()
    ok, lines = compare_grids(whole, parts)
    assert ok, lines
    hs.set_scan(scans[-1][1])
    hs.T = synth.pose_matrix(*scans[-1][0])
    sc = hs.scan()
    rays = hs.normalized_rays(cfg.cell_size).copy()
    c1, n1, m1, k1 = whole.raycast_mask(sc, rays)
    c2, n2, m2, k2 = parts.raycast_mask(sc, rays)
    assert same(m1, m2) and k1 == k2 and k1 > 0
    assert same(c1[m1 > 0], c2[m2 > 0]) and same(n1[m1 > 0], n2[m2 > 0])


def test_scan_box_contains_every_touched_partition():
    cfg = synth.config("C1")
    g = capi.Grid(cfg.cell_size, 5, cfg.layout_grid)
    g.set_max_truncation(cfg.max_truncation)
    hs = HostSensor(cfg.sensor, capi.invert3x3)
    pose, r = next(iter(cfg.scans(1)))
    hs.set_scan(r)
    hs.T = synth.pose_matrix(*pose)
    sc = hs.scan()
    before = g.partition_states()[0].copy()
    g.push(sc)
    st, iw = g.partition_states()
    px0, py0, px1, py1 = g.scan_box(sc)
    parts_x = (1 << cfg.layout_grid) // 32
    touched = np.nonzero(st != before)[0]
    assert len(touched) > 0
    assert (touched % parts_x >= px0).all() and (touched % parts_x <= px1).all()
    assert (touched // parts_x >= py0).all() and (touched // parts_x <= py1).all()


def test_batched_pushes_on_bands():
    """Two scans per launch pair (tsdg_push_batch) on a sharded grid == the same scans pushed one by one into the
    unsharded grid."""
    from ohm_tsd_slam_b200.workload import DoubleLaserWorkload
    wl = DoubleLaserWorkload("C1", n_map=4, n_steps=2, invert=capi.invert3x3)
    cfg = wl.cfg
    whole = capi.Grid(cfg.cell_size, 5, cfg.layout_grid)
    parts = LocalBands(cfg.cell_size, cfg.layout_grid, 3, peer=True)
    whole.set_max_truncation(cfg.max_truncation)
    parts.set_max_truncation(cfg.max_truncation)
    scans = wl.map_scans + [sc for st in wl.step_scans for sc in st]
    for k in range(0, len(scans), 2):
        for sc in scans[k:k + 2]:
            whole.push(sc)
        parts.push_batch(scans[k:k + 2], sync=(k % 4 == 2))
    parts.sync_halos()
    ok, lines = compare_grids(whole, parts)
    assert ok, lines


@pytest.mark.parametrize("name,bands", [("tiny", 2), ("C1", 1), ("C1", 3), ("C1", 7)])
def test_library_sharded_handle_equals_single_grid(name, bands):
    """tsdg_create_sharded (csrc/sharded.cu): the SLAM loop -- ray cast from the current pose, sample, push -- on ONE
    handle whose bands, halos, flags and ray-cast exchange live inside the library; every result equals the unsharded
    grid's bit for bit, reads synchronising lazily after the pushes."""
    cfg = synth.config(name)
    whole = capi.Grid(cfg.cell_size, 5, cfg.layout_grid)
    parts = capi.ShardedGrid(cfg.cell_size, 5, cfg.layout_grid, bands)
    assert parts.n_bands == bands
    whole.set_max_truncation(cfg.max_truncation)
    parts.set_max_truncation(cfg.max_truncation)
    hs = HostSensor(cfg.sensor, capi.invert3x3)
    scans = list(cfg.scans(8))
    (x, y, th), r0 = scans[0]
    assert whole.free_footprint(x, y, 0.6, 0.6) and parts.free_footprint(x, y, 0.6, 0.6)
    assert not whole.free_footprint(-50.0, -50.0, 0.6, 0.6) and not parts.free_footprint(-50.0, -50.0, 0.6, 0.6)
    rng = np.random.default_rng(5)
    side = cfg.cell_size * (1 << cfg.layout_grid)
    for k, (pose, r) in enumerate(scans):
        hs.set_scan(r)
        hs.T = synth.pose_matrix(*pose)
        sc = hs.scan()
        if k > 0:
            rays = hs.normalized_rays(cfg.cell_size).copy()
            c1, n1, m1, k1 = whole.raycast_mask(sc, rays)
            c2, n2, m2, k2 = parts.raycast_mask(sc, rays)
            assert same(m1, m2) and k1 == k2 and k1 > 0
            assert same(c1[m1 > 0], c2[m2 > 0]) and same(n1[m1 > 0], n2[m2 > 0])
            xy = np.concatenate([c1[m1 > 0] + rng.normal(0, 0.03, (k1, 2)), rng.uniform(-0.1 * side, 1.1 * side, (200, 2))])
            t1, s1 = whole.interpolate_bilinear(xy)
            t2, s2 = parts.interpolate_bilinear(xy)
            assert same(s1, s2) and same(t1[s1 == 0], t2[s2 == 0]) and (s1 == 0).sum() > 0
        if k % 3 == 2:
            whole.push(sc)
            whole.push(sc)
            parts.push_batch([sc, sc])
        else:
            whole.push(sc)
            parts.push(sc)
            a, b = whole.last_push_stats(), parts.last_push_stats()
            assert a["cell_updates"] == b["cell_updates"] and a["active_tiles"] == b["active_tiles"]
    ok, lines = compare_grids(whole, parts)
    assert ok, lines
    parts.close()
