"""Pins the plain-C port oracle (oracle/port) to the reference:
  * against the golden vectors generated from the reference's own compiled sources (tests/golden/), and
  * directly against oracle/_ref/libohm_ref.so where that prebuilt library is present.
CPU only."""
import math

import numpy as np
import pytest

from ohm_tsd_slam_b200 import synth
from ohm_tsd_slam_b200.scan import HostSensor
from oracle import port, ref
from tests.harness import GOLDEN, replay_golden_sequence, same


@pytest.mark.parametrize("name", ["tiny", "C1"])
def test_port_replays_reference_golden_bit_exact(name):
    fails = replay_golden_sequence(port, name, exact_icp=True)
    assert not fails, "\n".join(fails)


def _golden_matcher_inputs():
    G = np.load(f"{GOLDEN}/matchers_tiny.npz")
    cfg = synth.config("tiny")
    return G, cfg


def test_port_matchers_reproduce_reference_golden():
    """Full match() of the three matchers under the replayed rand() stream == the reference's result."""
    G, cfg = _golden_matcher_inputs()
    g = port.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
    g.set_max_truncation(cfg.max_truncation)
    icp = port.Icp(30, 0.4, 0.02, g.bounds)
    hs = HostSensor(cfg.sensor, port.invert3x3)
    scans = list(cfg.scans(4))
    (x, y, th), r0 = scans[0]
    hs.set_scan(r0)
    hs.transform(synth.pose_matrix(x, y, th))
    g.push(hs.scan())
    res = cfg.sensor.angular_res
    phimax = math.radians(30.0)
    for k, (_, r) in enumerate(scans[1:]):
        hs.set_scan(r)
        assert same(hs.pose, G[f"pose_{k}"])
        M, mM, S, mS = G[f"M_{k}"], G[f"maskM_{k}"], G[f"S_{k}"], G[f"maskS_{k}"]
        for trials, ctrl in ((20, 60), (100, 140)):
            tag = f"{k}_{trials}_{ctrl}"
            port.seed(1000 + k)
            assert same(port.match_tsd(g, trials, 0.15, ctrl, 0.25, hs.pose, M, mM, S, mS, phimax, 0.25, res), G[f"tsd_{tag}"])
            port.seed(2000 + k)
            assert same(port.match_rnm(trials, 0.15, ctrl, M, mM, S, mS, phimax, 0.25, res), G[f"rnm_{tag}"])
            port.seed(3000 + k)
            assert same(port.match_pdf(trials, 0.15, ctrl, ref.PDF_DEFAULTS, M, mM, S, mS, phimax, 0.25, res), G[f"pdf_{tag}"])
            assert not same(G[f"tsd_{tag}"], np.eye(3))
        sc = hs.scan()
        rays = hs.normalized_rays(cfg.cell_size).copy()
        c, n, m, _ = g.raycast_mask(sc, rays)
        scene, ms, _ = hs.scene()
        T = icp.run(c[m > 0], n[m > 0], scene[ms > 0], hs.pose)[0]
        hs.transform(T)
        g.push(hs.scan())


@pytest.mark.skipif(not ref.available(), reason="prebuilt reference library oracle/_ref/libohm_ref.so not present")
def test_port_equals_reference_library_live():
    """Same inputs through the reference's compiled sources and the port, bit for bit (push, raycast, ICP)."""
    cfg = synth.config("tiny")
    ref.set_threads(1)
    gr = ref.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
    gp = port.Grid(cfg.cell_size, cfg.layout_partition, cfg.layout_grid)
    gr.set_max_truncation(cfg.max_truncation)
    gp.set_max_truncation(cfg.max_truncation)
    sr = ref.Sensor(cfg.sensor)
    hs = HostSensor(cfg.sensor, ref.invert)
    assert same(ref.invert(synth.pose_matrix(3.1, 2.2, 0.3)), port.invert3x3(synth.pose_matrix(3.1, 2.2, 0.3)))
    icr = ref.Icp(30, 0.4, 0.02, gr.bounds)
    icp = port.Icp(30, 0.4, 0.02, gp.bounds)
    scans = list(cfg.scans(5))
    (x, y, th), r0 = scans[0]
    sr.set_scan(r0)
    hs.set_scan(r0)
    assert same(sr.mask, hs.mask) and same(sr.data, hs.data)
    sr.transform(synth.pose_matrix(x, y, th))
    hs.transform(synth.pose_matrix(x, y, th))
    gr.push(sr)
    gp.push(hs.scan())
    for _, r in scans[1:]:
        sr.set_scan(r)
        hs.set_scan(r)
        c1, n1, m1, _ = gr.raycast_mask(sr)
        rays = hs.normalized_rays(cfg.cell_size)
        assert same(rays, sr.normalized_rays(cfg.cell_size))
        c2, n2, m2, _ = gp.raycast_mask(hs.scan(), rays)
        assert same(m1, m2) and same(c1[m1 > 0], c2[m2 > 0]) and same(n1[m1 > 0], n2[m2 > 0])
        sc1, ms1, _ = sr.scene()
        sc2, ms2, _ = hs.scene()
        assert same(ms1, ms2) and same(sc1[ms1 > 0], sc2[ms2 > 0])
        a = icr.run(c1[m1 > 0], n1[m1 > 0], sc1[ms1 > 0], sr.pose)
        b = icp.run(c1[m1 > 0], n1[m1 > 0], sc1[ms1 > 0], hs.pose)
        assert same(a[0], b[0]) and a[1:] == b[1:]
        sr.transform(a[0])
        hs.transform(b[0])
        assert same(sr.pose, hs.pose)
        gr.push(sr)
        gp.push(hs.scan())
        s1, w1 = gr.partition_states()
        s2, w2 = gp.partition_states()
        assert same(s1, s2) and same(w1, w2)
        for p in np.nonzero(s1 == 2)[0]:
            ta, wa = gr.download_partition(int(p))
            tb, wb = gp.download_partition(int(p))
            assert same(ta, tb) and same(wa, wb)


def test_back_project_edge_cases():
    """backProject at the FOV edges and behind the sensor (-2 below, -1 above; SensorPolar2D.cpp:130-133)."""
    cfg = synth.config("tiny")
    hs = HostSensor(cfg.sensor, port.invert3x3)
    hs.set_scan(np.full(cfg.sensor.beams, 2.0, dtype=np.float32))
    hs.transform(synth.pose_matrix(3.0, 3.0, 0.4))
    sc = hs.scan()
    ang = np.array([cfg.sensor.phi_min - 0.01, cfg.sensor.phi_min, 0.0, cfg.sensor.phi_upper - 1e-9, cfg.sensor.phi_upper + 0.01,
                    math.pi, -math.pi + 1e-6]) + 0.4
    xy = np.stack([3.0 + 1.5 * np.cos(ang), 3.0 + 1.5 * np.sin(ang)], axis=1)
    idx = port.back_project(sc, xy)
    assert idx[0] == -2 and idx[1] == 0 and idx[2] == (cfg.sensor.beams - 1) // 2
    assert idx[3] == cfg.sensor.beams - 1 and idx[4] == -1 and idx[5] == -1 and idx[6] == -2


@pytest.mark.parametrize("name", ["tiny", "C1"])
def test_port_map_publication_matches_reference_golden(name):
    """RayCastAxisAligned2D::calcCoords + TsdGrid::grid2ColorImage of the port vs the fixtures
    generated from the reference."""
    from tests.harness import check_axis_map
    assert check_axis_map(port, name) == []


def test_port_checkpoint_format_matches_reference_library(tmp_path):
    """storeGrid / file constructor of the port vs the reference library, where it is present: identical bytes,
    identical grid after loading and after the next push."""
    if not ref.available():
        pytest.skip("oracle/_ref/libohm_ref.so not built here")
    from tests.harness import axis_map_scenario, compare_grids
    cfg = synth.config("tiny")
    gp = port.Grid(cfg.cell_size, 5, cfg.layout_grid)
    gr = ref.Grid(cfg.cell_size, 5, cfg.layout_grid)
    gp.set_max_truncation(cfg.max_truncation)
    gr.set_max_truncation(cfg.max_truncation)
    sen = ref.Sensor(cfg.sensor)
    scans = axis_map_scenario(cfg, 6, ref.invert)

    def push_ref(g, sc):
        sen.set_data(sc.ranges, sc.mask)
        sen.pose = sc.pose
        g.push(sen)

    for sc in scans[:5]:
        gp.push(sc)
        push_ref(gr, sc)
    fp, fr = str(tmp_path / "port.txt"), str(tmp_path / "ref.txt")
    assert gp.store(fp) and gr.store(fr)
    assert open(fp, "rb").read() == open(fr, "rb").read()
    lp, lr = port.Grid.load(fr), ref.Grid.load(fr)
    ok, lines = compare_grids(lr, lp)
    assert ok, lines
    lp.push(scans[5])
    push_ref(lr, scans[5])
    ok, lines = compare_grids(lr, lp)
    assert ok, lines
