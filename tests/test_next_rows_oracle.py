"""CPU: the oracle's restatements of the SURVEY 8(f) rows -- pcl::VoxelGrid (N4) and Map::AssociatePlanesByBoundary (N1) --
against hand-computed cases and an independent numpy restatement."""
import numpy as np

from oracle import pyoracle
from tests.next_util import POINT, make_points, np_associate, np_voxel_groups


def test_voxel_grid_known_answers():
    # two voxels at leaf 0.1: (0.01, 0.02, 0.03) + (0.05, 0.06, 0.07) share voxel (0,0,0); (0.31, 0.02, 0.03) sits in (3,0,0)
    pts = make_points([(0.31, 0.02, 0.03), (0.01, 0.02, 0.03), (0.05, 0.06, 0.07)],
                      [(255, 11, 0, 251), (255, 10, 0, 250), (255, 11, 0, 251)])
    out, idx = pyoracle.voxel_grid(pts, 0.1)
    assert len(out) == 2 and list(idx) == [0, 3]                       # ascending voxel index, not input order
    f = np.float32
    assert out[0]["x"] == (f(0.01) + f(0.05)) / f(2) and out[0]["z"] == (f(0.03) + f(0.07)) / f(2)
    assert out[0]["rgba"] == (255 << 24) | (10 << 16) | (0 << 8) | 250  # uint32(10.5), uint32(250.5): truncation
    assert out[1]["x"] == f(0.31) and out[1]["rgba"] == pts[0]["rgba"]
    # negative coordinates: floor, not truncation; index = ijk . (1, div_x, div_x * div_y)
    pts = make_points([(-0.05, 0.0, 0.0), (0.05, 0.0, 0.0), (0.05, 0.15, 0.0), (-0.05, 0.15, 0.25)])
    out, idx = pyoracle.voxel_grid(pts, 0.1)
    assert list(idx) == [0, 1, 3, 2 + 2 * 2 * 2]                        # div_b = (2, 2, 3)
    # a non-finite point is skipped; an empty cloud gives an empty cloud
    pts = make_points([(0.0, 0.0, 0.0), (np.nan, 0.0, 0.0), (0.0, 0.0, 0.01)])
    out, _ = pyoracle.voxel_grid(pts, 0.1)
    assert len(out) == 1 and out[0]["z"] == np.float32(0.01) / np.float32(2)
    assert len(pyoracle.voxel_grid(np.empty(0, POINT), 0.1)[0]) == 0


def test_voxel_grid_leaf_too_small_returns_the_input():
    pts = make_points([(0.0, 0.0, 0.0), (100.0, 100.0, 100.0), (50.0, 1.0, 2.0), (50.0, 1.0, 2.0)])
    out, idx = pyoracle.voxel_grid(pts, 1e-3)                           # (1e5)^3 voxels > INT_MAX
    assert np.array_equal(out, pts) and np.all(idx == -1)


def test_voxel_grid_against_numpy_grouping():
    rng = np.random.default_rng(7)
    xyz = (rng.normal(size=(6000, 3)) * [1.5, 0.8, 0.4] + [0.3, -0.2, 2.5]).astype(np.float32)
    rgba = rng.integers(0, 256, size=(6000, 4))
    pts = make_points(xyz, rgba)
    for leaf in (0.01, 0.05, (0.2, 0.1, 0.05)):
        out, idx = pyoracle.voxel_grid(pts, leaf)
        keys, members = np_voxel_groups(pts, leaf)
        assert np.array_equal(idx.astype(np.int64), keys)
        assert len(out) == len(keys)
        for k in range(0, len(keys), 37):
            m = pts[members[k]]
            n = len(m)
            for ax in "xyz":
                exact = float(np.sum(m[ax].astype(np.float64)) / n)
                assert abs(float(out[k][ax]) - exact) <= (n + 1) * 2.0 ** -23 * max(1.0, np.abs(m[ax]).max())
            r = int(np.float32(np.sum(((m["rgba"] >> 16) & 255).astype(np.float32))) / np.float32(n))
            assert (int(out[k]["rgba"]) >> 16) & 255 == r


def test_association_known_answers():
    ring = lambda z, r: make_points([(r * np.cos(t), r * np.sin(t), z) for t in np.linspace(0, 6.2, 40)])
    frame = np.array([[0, 0, 1, -2.0]], np.float32)                     # the plane z = 2
    map_w = np.array([[0, 0, 1, -2.5],       # parallel, boundary 0.5 m away: not associated (0.5 > 0.2) -> parallel candidate
                      [0, 0, -1, 2.05],      # same plane seen from the other side (angle -1), boundary 0.05 m away -> associated
                      [1, 0, 0, -1.0],       # perpendicular -> vertical
                      [0, 0, 1, -2.01],      # closer still (0.01 m) -> replaces the association
                      [0.0, 0.6, 0.8, -1.0]  # 37 degrees: neither
                      ], np.float32)
    bnds = [ring(2.5, 1.0), ring(2.05, 1.0), make_points([(1.0, y, 2.0) for y in np.linspace(-1, 1, 30)]), ring(2.01, 0.5), ring(1.0, 0.2)]
    a, v, p, d = pyoracle.associate_planes(frame, map_w, bnds)
    assert (a[0], v[0], p[0]) == (3, 2, 0)
    assert d[0] == abs(np.float32(2.01) + np.float32(-2.0))
    # an associated plane is not also a parallel candidate (`continue`), and the not-seen list is only consulted when
    # nothing in the seen list associated
    a, v, p, d = pyoracle.associate_planes(frame, map_w[[1, 3]], [bnds[1], bnds[3]], n_seen=1)
    assert (a[0], v[0], p[0]) == (0, -1, -1)
    a, v, p, d = pyoracle.associate_planes(frame, map_w[[0, 3]], [bnds[0], bnds[3]], n_seen=1)
    assert (a[0], p[0]) == (1, 0)
    # no map planes / empty boundary cloud: PointDistanceFromPlane returns 100
    a, v, p, d = pyoracle.associate_planes(frame, map_w[[1]], [np.empty(0, POINT)])
    assert a[0] == -1 and d[0] == np.float32(0.2)


def test_association_against_numpy_restatement():
    rng = np.random.default_rng(11)
    for trial in range(20):
        n_map, n_pl = int(rng.integers(1, 30)), int(rng.integers(1, 12))
        nrm = rng.normal(size=(n_map, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        base = np.eye(3)[rng.integers(0, 3, n_map)] * rng.choice([-1, 1], (n_map, 1))
        nrm = np.where(rng.random((n_map, 1)) < 0.7, base + 0.02 * rng.normal(size=(n_map, 3)), nrm)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        map_w = np.concatenate([nrm, rng.uniform(-3, 3, (n_map, 1))], 1).astype(np.float32)
        bnds = [make_points(rng.uniform(-3, 3, (int(rng.integers(0, 200)), 3))) for _ in range(n_map)]
        pn = np.eye(3)[rng.integers(0, 3, n_pl)] + 0.03 * rng.normal(size=(n_pl, 3)); pn /= np.linalg.norm(pn, axis=1, keepdims=True)
        plane_w = np.concatenate([pn, rng.uniform(-3, 3, (n_pl, 1))], 1).astype(np.float32)
        n_seen = int(rng.integers(0, n_map + 1))
        got = pyoracle.associate_planes(plane_w, map_w, bnds, n_seen=n_seen)
        ref = np_associate(plane_w, map_w, bnds, n_seen)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r)


def test_transform_cloud_known_answers():
    """pcl::transformPointCloud with a Matrix4d (MapPlane::UpdateBoundary): double arithmetic, one rounding to float"""
    pts = make_points([(1.0, 2.0, 3.0), (0.1, 0.2, 0.3), (-4.5, 0.0, 7.25)], [(255, 1, 2, 3)] * 3)
    eye = np.eye(4)
    assert np.array_equal(pyoracle.transform_cloud(pts, eye), pts)
    rz = np.array([[0, -1, 0, 0.5], [1, 0, 0, -1.0], [0, 0, 1, 2.0], [0, 0, 0, 1.0]])     # 90 degrees about z + translation
    out = pyoracle.transform_cloud(pts, rz)
    assert np.array_equal(out["x"], (-pts["y"].astype(np.float64) + 0.5).astype(np.float32))
    assert np.array_equal(out["y"], (pts["x"].astype(np.float64) - 1.0).astype(np.float32))
    assert np.array_equal(out["z"], (pts["z"].astype(np.float64) + 2.0).astype(np.float32))
    assert np.array_equal(out["rgba"], pts["rgba"])
    # the sum is formed in double, left to right: ((m00 x + m01 y) + m02 z) + m03
    rng = np.random.default_rng(2)
    m = np.eye(4); m[:3, :] = rng.normal(size=(3, 4))
    p = make_points(rng.normal(size=(500, 3)) * 3)
    out = pyoracle.transform_cloud(p, m)
    x, y, z = (p[a].astype(np.float64) for a in "xyz")
    for r, ax in enumerate("xyz"):
        ref = (((m[r, 0] * x + m[r, 1] * y) + m[r, 2] * z) + m[r, 3]).astype(np.float32)
        assert np.array_equal(out[ax], ref)
    assert len(pyoracle.transform_cloud(np.empty(0, POINT), eye)) == 0
