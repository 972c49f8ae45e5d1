"""CPU: known-answer tests that pin the oracle to the published PCL 1.8.0 / SP-SLAM behaviour.
The reference ships no tests for this path, so these are the anchors: closed-form scenes and stage-level KATs."""
import numpy as np
import pytest

from sp_slam_b200 import scenes


def test_chamfer_repeated_1p4_sums(oracle_lib):
    # a single edge pixel at the centre of a large mask: diagonal steps accumulate 1.4f in fp32 order
    m = np.full((41, 41), 255, np.uint8)
    m[20, 20] = 0
    d = oracle_lib.chamfer(m)
    diag = [d[20 + k, 20 + k] for k in range(1, 8)]
    expect = []
    v = np.float32(0)
    for _ in range(7):
        v = np.float32(v + np.float32(1.4))
        expect.append(v)
    assert [float(x) for x in diag] == [float(x) for x in expect]
    assert float(diag[4]) == 7.0 and float(diag[2]) == float(np.float32(4.1999998))
    # axial steps cost 1.0, and the map is the min over 8-neighbour chamfer paths
    assert float(d[20, 25]) == 5.0 and float(d[17, 20]) == 3.0
    assert float(d[20 + 1, 20 + 3]) == float(np.float32(np.float32(1.4) + np.float32(2.0)))
    # no edge anywhere: every pixel keeps width + height
    d2 = oracle_lib.chamfer(np.full((12, 17), 255, np.uint8))
    assert np.all(d2 == 29.0)


def test_chamfer_row_wraparound(oracle_lib):
    # PCL's forward pass reads previous_row[w] (= current_row[0]) at the last column: an edge at column 0 of row r
    # reaches (r, w-1) with cost 1.4 although the pixels are w-1 apart
    m = np.full((8, 30), 255, np.uint8)
    m[4, 0] = 0
    d = oracle_lib.chamfer(m)
    assert float(d[4, 29]) == float(np.float32(1.4))
    # the alternative reading (no aliasing across the row end) is a live switch of the oracle: the same pixel is then far away
    assert float(oracle_lib.chamfer(m, alt=oracle_lib.ALT_CHAMFER_NO_WRAP)[4, 29]) > 10.0


def test_alternative_readings_of_the_line_fit(oracle_lib):
    # ORC_ALT_SAMPLE_GOOD_OR / ORC_ALT_RNG_MASK: isSampleGood with && rejects a pair that shares one
    # coordinate, so on points of a line with constant z the default reading finds no sample at all; || accepts them
    t = np.linspace(0.0, 1.0, 200).astype(np.float32)
    pts = np.zeros(200, oracle_lib.POINT_DTYPE)
    pts["x"], pts["y"], pts["z"] = t, 2.0 * t, 1.5
    coef, inl, it = oracle_lib.sac_line(pts)
    assert len(inl) == 0                                         # PCL 1.8.0: "no samples could be selected"
    coef, inl, it = oracle_lib.sac_line(pts, alt=oracle_lib.ALT_SAMPLE_GOOD_OR)
    assert len(inl) == 200 and abs(abs(coef[3] / coef[4]) - 0.5) < 1e-5
    # the two RNG readings draw different samples (mt() >> 1 against mt() & INT_MAX) but fit the same line here
    rng = np.random.default_rng(3)
    pts["z"] = 1.5 + 0.3 * t + rng.normal(0, 1e-4, 200).astype(np.float32)
    a = oracle_lib.sac_line(pts)
    b = oracle_lib.sac_line(pts, alt=oracle_lib.ALT_RNG_MASK)
    assert len(a[1]) > 150 and len(b[1]) > 150


def test_is_border_point_without_a_finite_projection(oracle_lib):
    # a line point at depth 0 projects to u = fx * 0 * inf + cx = NaN: the reference's window loops never run, res / num is
    # NaN, `PcZ - NaN > 0.1` is false and IsBorderPoint returns TRUE (src/Frame.cc:1026-1056); a finite point far behind the
    # depth image is not a border point
    depth = np.full((480, 640), 2.0, np.float32)
    assert oracle_lib.is_border_point(depth, 0.0, 0.0, 0.0)
    assert oracle_lib.is_border_point(depth, 0.1, 0.1, 2.0)              # lies on the surface
    assert not oracle_lib.is_border_point(depth, 0.1, 0.1, 2.5)          # occluded: 0.5 m behind the measured depth
    assert not oracle_lib.is_border_point(depth, 0.1, 0.1, -1.0)         # PcZ < 0


def test_eigen33_smallest_known(oracle_lib):
    # diagonal matrix: eigenvalues are the entries, eigenvector of the smallest is its axis
    ev, vec = oracle_lib.eigen33_smallest(np.diag([3.0, 0.5, 2.0]))
    assert abs(ev - 0.5) < 1e-6 and abs(abs(vec[1]) - 1.0) < 1e-6
    # rank-2 covariance of points on the plane x + 2y + 2z = 0: normal (1,2,2)/3, eigenvalue 0 (quadratic branch)
    rng = np.random.default_rng(7)
    a = np.array([2.0, -1.0, 0.0]) / np.sqrt(5.0)
    n = np.array([1.0, 2.0, 2.0]) / 3.0
    b = np.cross(n, a)
    p = rng.normal(size=(4000, 1)) * a + rng.normal(size=(4000, 1)) * 0.5 * b
    cov = np.cov(p.T, bias=True)
    ev, vec = oracle_lib.eigen33_smallest(cov)
    assert abs(ev) < 1e-6
    assert abs(abs(float(vec @ n)) - 1.0) < 1e-5
    evals, big = oracle_lib.eigen33_largest(cov)
    w = np.linalg.eigvalsh(cov)
    assert np.allclose(evals, w, atol=1e-5)
    assert abs(abs(float(big @ a)) - 1.0) < 1e-4


def test_ransac_draw_sequence_is_mt19937_12345(oracle_lib):
    # drawIndexSample: swap(sh[i], sh[i + rnd() % (n - i)]) for i = 0, 1 with rnd() = mt19937(12345)() >> 1
    n = 1000
    draws = oracle_lib.ransac_draws(n, 64)
    mt = np.random.MT19937()
    # numpy's legacy seeding (init_genrand) is the same recurrence as std::mt19937::seed(uint32)
    mt._legacy_seeding(12345)
    raw = mt.random_raw(128)
    sh = list(range(n))
    for k in range(64):
        for i in range(2):
            j = i + int(raw[2 * k + i] >> 1) % (n - i)
            sh[i], sh[j] = sh[j], sh[i]
        assert (sh[0], sh[1]) == tuple(int(x) for x in draws[k])


def test_mt19937_reference_value():
    # the 10000th output of a default-seeded (5489) mt19937 is 4123659995 (C++11 [rand.predef])
    mt = np.random.MT19937()
    mt._legacy_seeding(5489)
    assert int(mt.random_raw(10000)[-1]) == 4123659995


def _plane_depth(normal, d, intr, rows, cols):
    """z-depth image of the plane n.p + d = 0 seen by the pinhole camera"""
    u, v = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    rx, ry = (u - intr.cx) / intr.fx, (v - intr.cy) / intr.fy
    z = -d / (normal[0] * rx + normal[1] * ry + normal[2])
    return z.astype(np.float32)


@pytest.mark.parametrize("normal,d", [((0.0, 0.0, -1.0), 2.0), ((0.3, -0.2, -0.93), 2.5), ((-0.5, 0.1, -0.86), 1.8)])
def test_single_analytic_plane(oracle_lib, normal, d):
    n = np.array(normal, float)
    n /= np.linalg.norm(n)
    depth = _plane_depth(n, d, scenes.TUM1, 480, 640)
    o = oracle_lib.Oracle().run(depth)
    assert o.width == 214 and o.height == 160
    assert o.n_real == 1
    pl = o.planes()[0]
    c = pl["coef"].astype(np.float64)
    assert c[3] >= 0                                                  # src/Frame.cc:918-919
    sign = 1.0 if np.dot(c[:3], n) > 0 else -1.0
    ang = np.arctan2(np.linalg.norm(np.cross(c[:3], sign * n)), np.dot(c[:3], sign * n))
    assert ang < 1e-4 and abs(c[3] - abs(d)) < 1e-4
    # the 10-pixel border has NaN normals, so the segment is the interior; refine() grows it back into the border
    nrm = o.normals()
    valid = np.isfinite(nrm[0]).reshape(160, 214)
    assert not valid[:10].any() and not valid[-10:].any() and not valid[:, :10].any() and not valid[:, -10:].any()
    assert valid[10:-10, 10:-10].all()
    m = o.models()[0]
    assert m["n_segment"] == (160 - 20) * (214 - 20)
    assert len(m["inliers"]) > m["n_segment"]
    # every NaN-normal pixel is its own component: label of the big component = number of singletons before it
    lab, n_lists = o.labels_raw()
    assert m["label"] == 10 * 214 + 10
    assert n_lists == (160 * 214 - m["n_segment"]) + 1 + 1
    # refine() grows the plane over the whole image, so the start pixel has no foreign in-image neighbour: PCL returns
    # an EMPTY contour and SP-SLAM substitutes every 20th inlier with a default-constructed colour (src/Frame.cc:1001-1011)
    # (the bottom-right corner is the one pixel neither raster pass can claim: pass 1 claims right/down only from rows
    # <= h-2 and columns <= w-2, pass 2 claims left/up)
    assert len(m["inliers"]) == 160 * 214 - 1 and 160 * 214 - 1 not in m["inliers"] and len(m["contour"]) == 0
    assert len(pl["boundary"]) == (len(pl["points"]) + 19) // 20
    assert np.array_equal(pl["boundary"]["x"], pl["points"]["x"][::20]) and np.all(pl["boundary"]["rgba"] == 0xFF000000)
    assert np.all(pl["points"]["rgba"] == 0xFF0000FA)                   # (r, g, b) = (0, 0, 250), src/Frame.cc:866-868


def test_contour_starts_at_last_inlier_and_is_closed(oracle_lib):
    depth = scenes.render(scenes.boxroom_rects(), scenes.poses(1000)[[200]], scenes.TUM1)[0]
    o = oracle_lib.Oracle(enable_supposed=0).run(depth)
    seen = 0
    for m in o.models():
        if len(m["contour"]):
            assert m["contour"][0] == m["inliers"][-1] == m["contour"][-1]
            lab = o.labels_refined().ravel()
            assert np.all(lab[m["contour"]] == lab[m["inliers"][0]])
            seen += 1
    assert seen >= 2


def test_min_inliers_is_strict_and_curvature_filter(oracle_lib):
    n = np.array([0.0, 0.0, -1.0])
    depth = _plane_depth(n, 2.0, scenes.TUM1, 480, 640)
    seg = (160 - 20) * (214 - 20)
    assert oracle_lib.Oracle(min_size=seg).run(depth).n_real == 0       # size > min_inliers, not >=
    assert oracle_lib.Oracle(min_size=seg - 1).run(depth).n_real == 1


def test_cloud_subsampling_and_backprojection(oracle_lib):
    rng = np.random.default_rng(3)
    depth = rng.uniform(0.5, 4.0, size=(480, 640)).astype(np.float32)
    o = oracle_lib.Oracle(enable_supposed=0).run(depth)
    x, y, z = o.cloud()
    it = scenes.TUM1
    fx, fy, cx, cy = (np.float32(v) for v in (it.fx, it.fy, it.cx, it.cy))
    for (r, c) in [(0, 0), (5, 7), (159, 213), (80, 100)]:
        i = r * 214 + c
        zz = depth[3 * r, 3 * c]
        assert z[i] == zz
        assert x[i] == np.float32(np.float32(np.float32(3 * c) - cx) * zz) / fx     # ((n - cx) * z) / fx
        assert y[i] == np.float32(np.float32(np.float32(3 * r) - cy) * zz) / fy
    assert o.n_real == 0   # white-noise depth has no plane


def test_supposed_plane_grid_is_50x50(oracle_lib):
    depth = scenes.render(scenes.boxroom_rects(), scenes.poses(1000)[[200]], scenes.TUM1)[0]
    o = oracle_lib.Oracle().run(depth)
    assert o.n_planes > o.n_real >= 1
    pls = o.planes()
    recs = [r for r in o.line_recs() if r["emitted"]]
    assert len(recs) == o.n_planes - o.n_real
    for k, rec in zip(range(o.n_real, o.n_planes), recs):
        sp = pls[k]
        assert len(sp["points"]) == 2500 + rec["n_inliers"]            # src/Frame.cc:1095-1106 + :988
        assert len(sp["boundary"]) == rec["n_inliers"]
        assert np.all(sp["points"]["rgba"][:2500] == 0xFF00FF00)       # green grid
        assert np.all(sp["points"]["rgba"][2500:] == 0xFFFF0000)       # red line points
        parent = pls[sp["src"]]["coef"]
        assert abs(float(np.dot(parent[:3], sp["coef"][:3]))) < 1e-4   # perpendicular to its parent
        assert sp["coef"][3] >= 0
    # de-duplication rule (src/Frame.cc:1116-1144): no two kept planes within 0.2 m and cos 20 deg
    for i in range(len(pls)):
        for j in range(i):
            a, b = pls[i]["coef"], pls[j]["coef"]
            assert not (abs(a[3] - b[3]) <= 0.2 and abs(float(np.dot(a[:3], b[:3]))) >= 0.9397)


def test_sat_sums_are_exact_for_depth_data(oracle_lib):
    # the claim the CUDA tiles rely on: fp64 sums of fp32 central differences never round on depth-camera data
    P = scenes.poses(1000)
    for f, noisy in ((200, False), (640, True)):
        d = scenes.render(scenes.boxroom_rects(), P[[f]], scenes.TUM1)[0]
        if noisy:
            d = scenes.add_noise(d, f)
        assert oracle_lib.Oracle(enable_supposed=0).run(d).sat_exact()


def test_covariance_matrix_method_on_an_analytic_plane(oracle_lib):
    """IntegralImageNormalEstimation::COVARIANCE_MATRIX (normal_method = 1, the method BASELINE.json's north_star words; the
    reference selects AVERAGE_3D_GRADIENT): on an exact plane the smallest eigenvector is the plane normal and the curvature
    is ~0; the 10-pixel NaN border and the NaN curvature of the gradient method are as in PCL."""
    it = scenes.TUM1
    n = np.array([0.2, -0.1, -1.0]); n /= np.linalg.norm(n)
    v, u = np.mgrid[0:480, 0:640].astype(np.float64)
    rays = np.stack([(u - it.cx) / it.fx, (v - it.cy) / it.fy, np.ones_like(u)], -1)
    depth = (-2.5 / (rays @ n)).astype(np.float32)          # plane n . p + 2.5 = 0
    o = oracle_lib.Oracle(normal_method=1, enable_supposed=0).run(depth)
    nrm, cv = o.normals(), o.curvature()
    ok = ~np.isnan(nrm[0])
    assert ok.sum() == (214 - 20) * (160 - 20)
    ang = np.arccos(np.clip(np.abs(nrm[:, ok].T @ n), 0, 1))
    assert ang.max() < 2e-3 and np.median(ang) < 2e-4
    assert np.all(cv[ok] >= 0) and np.median(cv[ok]) < 1e-4
    assert o.n_real == 1
    g = oracle_lib.Oracle(enable_supposed=0).run(depth)
    assert np.all(np.isnan(g.curvature()))                 # AVERAGE_3D_GRADIENT never sets the curvature
    # the two recalled details of this method are switches as well
    a = oracle_lib.Oracle(normal_method=1, enable_supposed=0, alt=oracle_lib.ALT_COV_TRACE).run(depth).curvature()
    assert not np.array_equal(a[ok], cv[ok])
