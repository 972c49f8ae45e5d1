"""-m gpu: the SURVEY 8(f) rows built so far -- VoxelGrid (N4) and plane association (N1) -- through the C ABI against the oracle."""
import numpy as np
import pytest

from sp_slam_b200 import api, scenes
from tests.next_util import POINT, make_points, np_voxel_groups

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frames():
    d = scenes.boxroom_sequence(6, start=195)
    ext = api.PlaneExtractor(max_frames=6)
    res = ext.extract_batch(d)
    yield ext, d, [res.frame(k) for k in range(6)]
    ext.close()


def check_cloud(got, pts, leaf, oracle_lib):
    ref, idx = oracle_lib.voxel_grid(pts, leaf)
    assert len(got) == len(ref)
    if len(ref) == 0:
        return 0
    assert np.array_equal(got["rgba"], ref["rgba"])                     # colour sums are small integers: exact in any order
    keys, members = np_voxel_groups(pts, leaf) if idx[0] >= 0 else (None, None)
    if keys is None:
        assert np.array_equal(got, ref)                                 # "leaf too small": the input comes back
        return 0
    pop = np.array([len(m) for m in members])
    small = pop <= 2                                                    # a + b == b + a: identical to std::sort's order
    for ax in "xyz":
        assert np.array_equal(got[ax][small], ref[ax][small])
        big = ~small
        if big.any():                                                   # PCL's order inside a voxel is introsort's: n-term fp32 sums
            scale = np.array([np.abs(pts[ax][m]).max() for m, b in zip(members, big) if b])
            assert np.all(np.abs(got[ax][big].astype(np.float64) - ref[ax][big]) <= pop[big] * 2.0 ** -23 * np.maximum(scale, 1e-3))
    return int((~small).sum())


def test_voxel_grid_on_plane_clouds(frames, oracle_lib):
    ext, _, fps = frames
    clouds = [c for fp in fps for c in fp.mvPlanePoints] + [c for fp in fps for c in fp.mvBoundaryPoints]
    assert len(clouds) > 20
    n_big = 0
    for leaf in (0.01, 0.05, (0.2, 0.1, 0.05)):
        outs = ext.voxel_grid(clouds, leaf)
        assert len(outs) == len(clouds)
        for got, pts in zip(outs, clouds):
            n_big += check_cloud(got, pts, leaf, oracle_lib)
    assert n_big > 100                                                  # the tolerance branch was exercised


def test_voxel_grid_merged_world_cloud_and_edge_cases(frames, oracle_lib):
    """the drawers' use (src/MapDrawer.cc:100-116): clouds of several frames merged, leaf 0.01 / 0.02; plus empty clouds,
    non-finite points, one point, and a cloud whose leaf is too small for int indices"""
    ext, _, fps = frames
    merged = np.concatenate([fp.mvPlanePoints[0] for fp in fps])
    rng = np.random.default_rng(3)
    holes = make_points(rng.uniform(-1, 1, (300, 3)))
    holes["x"][::7] = np.nan
    holes["z"][5::11] = np.inf
    far = make_points([(0, 0, 0), (900.0, 900.0, 900.0), (1.0, 2.0, 3.0), (1.0, 2.0, 3.0)])
    clouds = [merged, np.empty(0, POINT), holes, make_points([(0.5, 0.25, 2.0)]), far, np.empty(0, POINT)]
    for leaf in (0.01, 0.02):
        outs = ext.voxel_grid(clouds, leaf)
        for got, pts in zip(outs, clouds):
            check_cloud(got, pts, leaf, oracle_lib)
    assert len(outs[1]) == 0 and len(outs[3]) == 1 and np.array_equal(outs[4], far)
    assert len(ext.voxel_grid([], 0.01)) == 0


def test_voxel_downsample_of_the_device_results(frames, oracle_lib):
    import torch
    ext, d, fps = frames
    dev = torch.from_numpy(d).cuda()
    for which, leaf in ((1, 0.05), (0, 0.03)):
        ext.extract_device(dev.data_ptr(), len(d), 480, 640)
        ext.voxel_downsample_results(leaf, which)
        res = ext.fetch()
        for k, fp in enumerate(fps):
            got = res.frame(k)
            assert got.mnPlaneNum == fp.mnPlaneNum and np.array_equal(got.mvPlaneCoefficients.view(np.uint32), fp.mvPlaneCoefficients.view(np.uint32))
            for i in range(fp.mnPlaneNum):
                raw, keep = (fp.mvBoundaryPoints, fp.mvPlanePoints) if which == 1 else (fp.mvPlanePoints, fp.mvBoundaryPoints)
                down, same = (got.mvBoundaryPoints, got.mvPlanePoints) if which == 1 else (got.mvPlanePoints, got.mvBoundaryPoints)
                assert np.array_equal(same[i], keep[i])
                check_cloud(down[i], raw[i], leaf, oracle_lib)
                assert len(down[i]) < len(raw[i]) or len(raw[i]) < 3


def test_plane_association_against_the_oracle(frames, oracle_lib):
    """Map planes = the planes of earlier frames (boundary clouds in their camera frame, which is the world frame of a
    camera that has not moved), frame planes = a later frame's; plus random maps with empty clouds and not-seen planes."""
    ext, _, fps = frames
    pm = api.PlaneMap(ext)
    map_w = np.concatenate([fp.mvPlaneCoefficients for fp in fps[:4]])
    bnds = [b for fp in fps[:4] for b in fp.mvBoundaryPoints]
    for n_seen in (len(map_w), len(map_w) // 2, 0):
        pm.upload(map_w, bnds, n_seen)
        for fp in fps[3:]:
            got = pm.associate(fp.mvPlaneCoefficients)
            ref = oracle_lib.associate_planes(fp.mvPlaneCoefficients, map_w, bnds, n_seen)
            for g, r in zip(got, ref):
                assert np.array_equal(g, r)
            assert (got[0] >= 0).any()
    rng = np.random.default_rng(5)
    for trial in range(10):
        n_map, n_pl = int(rng.integers(1, 60)), int(rng.integers(1, 40))
        nrm = np.eye(3)[rng.integers(0, 3, n_map)] * rng.choice([-1, 1], (n_map, 1)) + 0.03 * rng.normal(size=(n_map, 3))
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        mw = np.concatenate([nrm, rng.uniform(-3, 3, (n_map, 1))], 1).astype(np.float32)
        bb = [make_points(rng.uniform(-3, 3, (int(rng.integers(0, 3000)), 3))) for _ in range(n_map)]
        pn = np.eye(3)[rng.integers(0, 3, n_pl)] + 0.03 * rng.normal(size=(n_pl, 3)); pn /= np.linalg.norm(pn, axis=1, keepdims=True)
        pw = np.concatenate([pn, rng.uniform(-3, 3, (n_pl, 1))], 1).astype(np.float32)
        n_seen = int(rng.integers(0, n_map + 1))
        pm.upload(mw, bb, n_seen)
        got = pm.associate(pw)
        ref = oracle_lib.associate_planes(pw, mw, bb, n_seen)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r)
    pm.upload(np.empty((0, 4), np.float32), [], 0)
    a, v, p, dd = pm.associate(fps[0].mvPlaneCoefficients)
    assert np.all(a == -1) and np.all(dd == np.float32(0.2))
    pm.close()


def test_map_boundary_updates_against_the_oracle(frames, oracle_lib):
    """MapPlane::UpdateBoundary on the device: the transformed cloud replaces map plane j's boundary -- from a host cloud
    and from the device results of the last extract (host-input path and device-resident path) -- bit-identical to
    pcl::transformPointCloud as restated by the oracle; associations afterwards match the oracle on the updated map."""
    import torch
    ext, d, fps = frames
    rng = np.random.default_rng(9)
    pm = api.PlaneMap(ext)
    map_w = np.concatenate([fp.mvPlaneCoefficients for fp in fps[:2]])
    bnds = [b.copy() for fp in fps[:2] for b in fp.mvBoundaryPoints]
    pm.upload(map_w, bnds)

    def pose():
        a = rng.normal(size=3) * 0.05
        R = np.array([[1, -a[2], a[1]], [a[2], 1, -a[0]], [-a[1], a[0], 1]])
        u, _, vt = np.linalg.svd(R)
        T = np.eye(4); T[:3, :3] = u @ vt; T[:3, 3] = rng.normal(size=3) * 0.1
        return T

    # (1) host clouds, including one much larger than the plane's slot (moves to the end of the arena) and an empty one
    big = np.concatenate([b for fp in fps for b in fp.mvBoundaryPoints])
    for j, cloud in ((0, fps[4].mvBoundaryPoints[0]), (1, big), (2 % len(bnds), np.empty(0, api.POINT_DTYPE)), (0, fps[5].mvBoundaryPoints[0])):
        T = pose()
        pm.update_boundary(j, T, cloud)
        bnds[j] = oracle_lib.transform_cloud(cloud, T)
        assert np.array_equal(pm.boundary(j), bnds[j])
    for k in range(len(bnds)):
        assert np.array_equal(pm.boundary(k), bnds[k])                   # the arena rebuild kept every other plane
    # (2) from the device results: host-input extract, then device-resident extract
    res = ext.extract_batch(d)
    dev = torch.from_numpy(d).cuda()
    for source in ("host", "device"):
        if source == "device":
            ext.extract_device(dev.data_ptr(), len(d), 480, 640)
        for frame in (1, 4):
            fp = res.frame(frame)
            for plane in range(fp.mnPlaneNum):
                j = int(rng.integers(len(bnds)))
                T = pose()
                pm.update_boundary_from_result(j, T, frame, plane, len(fp.mvBoundaryPoints[plane]))
                bnds[j] = oracle_lib.transform_cloud(fp.mvBoundaryPoints[plane], T)
                assert np.array_equal(pm.boundary(j), bnds[j])
    with pytest.raises(api.SpxError):
        pm.update_boundary_from_result(0, np.eye(4), 1, 0, len(res.frame(1).mvBoundaryPoints[0]) + 1)   # size mismatch is refused
    with pytest.raises(api.SpxError):
        pm.update_boundary_from_result(1, np.eye(4), 1, 0, 100000)        # ... also when the claimed cloud outgrows the slot
    for k in range(len(bnds)):
        assert np.array_equal(pm.boundary(k), bnds[k])                   # ... and leaves the map as it was
    # (3) world coefficients + association on the updated map
    map_w = map_w.copy()
    map_w[0] = fps[3].mvPlaneCoefficients[0]
    pm.set_world_pos(0, map_w[0])
    for fp in fps[2:]:
        got = pm.associate(fp.mvPlaneCoefficients)
        ref = oracle_lib.associate_planes(fp.mvPlaneCoefficients, map_w, bnds)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r)
    pm.close()
