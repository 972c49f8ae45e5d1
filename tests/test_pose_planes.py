"""CPU: the plane part of Optimizer::PoseOptimization (host code in sp_slam_b200/host/PlanePoseOptimizer.h,
reached through the C ABI) against an independent numpy restatement of g2oAddition's residuals and a scipy solve.
Parity with g2o itself is unpinned (it cannot be built here); tolerances are stated per test."""
import numpy as np
import pytest
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation

from sp_slam_b200 import api


# ---- numpy restatement of g2oAddition/Plane3D.h (written from the header, independent of the C++ code) ----
def normalize(v):
    v = np.asarray(v, np.float64) / np.linalg.norm(v[:3])
    return -v if v[3] < 0 else v


def azimuth(v):
    return np.arctan2(v[1], v[0])


def elevation(v):
    return np.arctan2(v[2], np.hypot(v[0], v[1]))


def rotation(v):
    return (Rotation.from_rotvec([0, 0, azimuth(v)]) * Rotation.from_rotvec([0, -elevation(v), 0])).as_matrix()


def transform_plane(T, p):                      # Plane3D operator*(Isometry3D, Plane3D)
    n = T[:3, :3] @ p[:3]
    v = np.array([*n, p[3] - T[:3, 3] @ n])
    return normalize(-v if v[3] < 0 else v)


def residual(kind, T, plane_w, meas):
    loc = transform_plane(T, normalize(plane_w))
    m = normalize(meas)
    if kind == 0:
        n = rotation(loc[:3]).T @ m[:3]
        return np.array([azimuth(n), elevation(n), -loc[3] - (-m[3])])
    if kind == 1:
        nor = loc[:3] if m[:3] @ loc[:3] >= 0 else -loc[:3]
        n = rotation(nor).T @ m[:3]
        return np.array([azimuth(n), elevation(n), 0.0])
    v = np.cross(loc[:3], m[:3])
    b = Rotation.from_rotvec(np.pi / 2 * v / np.linalg.norm(v)).as_matrix() @ loc[:3]
    n = rotation(b).T @ m[:3]
    return np.array([azimuth(n), elevation(n), 0.0])


def pose(rotvec, t):
    T = np.eye(4)
    T[:3, :3] = Rotation.from_rotvec(rotvec).as_matrix()
    T[:3, 3] = t
    return T


ROOM = np.array([[0, 1, 0, 1.4], [0, -1, 0, 1.6], [1, 0, 0, 3.0], [-1, 0, 0, 3.0], [0, 0, 1, 2.5], [0, 0, -1, 2.5],
                 [0.6, 0, 0.8, 1.0], [0, 0.6, -0.8, 2.0]], np.float64)


def measurements(T, planes):
    return np.array([transform_plane(T, normalize(p)) for p in planes])


def test_struct_layout_and_residuals_against_numpy():
    assert api.EDGE_DTYPE.itemsize == 80
    rng = np.random.default_rng(4)
    T = pose(rng.normal(size=3) * 0.3, rng.normal(size=3))
    planes = np.concatenate([ROOM, np.c_[rng.normal(size=(20, 3)), rng.uniform(0.5, 4, 20)]])
    planes = np.array([normalize(p) for p in planes]).astype(np.float32)
    meas = measurements(pose(rng.normal(size=3) * 0.3, rng.normal(size=3)), planes).astype(np.float32)
    for kind in (0, 1, 2):
        e = api.plane_edges([kind] * len(planes), planes, meas)
        got = api.plane_edge_errors(T, e)
        ref = np.array([residual(kind, T, p.astype(np.float64), m.astype(np.float64)) for p, m in zip(planes, meas)])
        assert np.allclose(got, ref, rtol=0, atol=1e-12), kind
    # a plane seen from the pose it was measured at has zero error
    e = api.plane_edges([0] * len(ROOM), ROOM, measurements(T, ROOM))
    assert np.abs(api.plane_edge_errors(T, e)).max() < 1e-6           # (float32 coefficients)


@pytest.mark.parametrize("seed", range(6))
def test_recovers_the_pose_from_exact_plane_matches(seed):
    rng = np.random.default_rng(20 + seed)
    T_gt = pose(rng.normal(size=3) * 0.2, rng.normal(size=3) * 0.5)
    meas = measurements(T_gt, ROOM)
    T0 = pose(Rotation.from_matrix(T_gt[:3, :3]).as_rotvec() + rng.normal(size=3) * 0.03, T_gt[:3, 3] + rng.normal(size=3) * 0.05)
    e = api.plane_edges([0] * len(ROOM), ROOM, meas)
    T, outlier, chi2, bad = api.pose_optimize_planes(T0, e)
    assert bad == 0 and not outlier.any()
    assert np.abs(T - T_gt).max() < 2e-6                                  # float32 coefficients in, double arithmetic inside
    assert chi2.max() < 1e-6


def test_outlier_rounds_and_agreement_with_scipy():
    """12 plane edges (>= 10, so all four rounds run) with measurement noise, one gross outlier and a few parallel /
    vertical edges: the outlier is flagged, and the pose agrees with scipy's least-squares solution on the inlier
    residuals weighted by the same information (the kernels are off after the third round and every inlier is below the
    Huber threshold, so both minimise the same cost)."""
    rng = np.random.default_rng(77)
    T_gt = pose([0.05, -0.1, 0.02], [0.2, -0.1, 0.3])
    planes = np.concatenate([ROOM, [[0.36, 0.48, 0.8, 1.5], [0.8, -0.6, 0, 2.2]]])
    meas = measurements(T_gt, planes)
    meas[:, :3] += rng.normal(size=(len(planes), 3)) * 0.002
    meas[:, 3] += rng.normal(size=len(planes)) * 0.005
    meas = np.array([normalize(m) for m in meas])
    meas[4] = normalize([0.3, 0.1, 0.95, 2.9])                           # a wrong association
    kinds = [0] * len(planes) + [1, 2]
    pw = np.concatenate([planes, [planes[0] * [1, 1, 1, 2.0]], [planes[2]]])       # a plane parallel to the floor, a wall vertical to ...
    ms = np.concatenate([meas, [meas[0]], [meas[0]]])                              # ... the measured floor
    e = api.plane_edges(kinds, pw, ms)
    T0 = pose([0.03, -0.07, 0.0], [0.1, 0.0, 0.2])
    T, outlier, chi2, bad = api.pose_optimize_planes(T0, e)
    assert bad == 1 and list(np.nonzero(outlier)[0]) == [4] and chi2[4] > 300.0
    inl = [i for i in range(len(kinds)) if i != 4]
    w = np.sqrt(e["info"])

    def fun(x):
        Tx = pose(x[:3], x[3:])
        return np.concatenate([(residual(kinds[i], Tx, e["plane_w"][i].astype(np.float64), e["measurement"][i].astype(np.float64)) * w[i])
                               for i in inl])
    sol = least_squares(fun, np.r_[Rotation.from_matrix(T0[:3, :3]).as_rotvec(), T0[:3, 3]], xtol=1e-14, ftol=1e-14, gtol=1e-14)
    T_ref = pose(sol.x[:3], sol.x[3:])
    assert np.abs(T - T_ref).max() < 5e-6
    assert np.abs(T - T_gt).max() < 0.02                                  # and both sit near the truth


def test_few_edges_stop_after_the_first_round_and_bad_arguments():
    """`if(optimizer.edges().size()<10) break;` (src/Optimizer.cc:1144): with fewer than ten edges only one round runs, so a
    gross outlier is flagged but never excluded from the solve"""
    T_gt = pose([0.0, 0.05, 0.0], [0.1, 0.0, 0.0])
    meas = measurements(T_gt, ROOM[:6])
    meas[1] = normalize([0.5, -0.8, 0.1, 0.4])
    e = api.plane_edges([0] * 6, ROOM[:6], meas)
    T, outlier, chi2, bad = api.pose_optimize_planes(np.eye(4), e)
    assert bad >= 1 and outlier[1]
    T4, *_ = api.pose_optimize_planes(np.eye(4), e, rounds=1)
    assert np.array_equal(T, T4)
    with pytest.raises(api.SpxError):
        bad_e = e.copy(); bad_e["kind"][0] = 7
        api.pose_optimize_planes(np.eye(4), bad_e)
    T_same, outl, _, nb = api.pose_optimize_planes(T_gt, np.zeros(0, api.EDGE_DTYPE))
    assert np.allclose(T_same, T_gt, atol=1e-15) and nb == 0


def test_cpp_header_stand_alone_with_extra_terms(tmp_path):
    """sp_slam_b200/host/PlanePoseOptimizer.h used directly from C++ (no CUDA, no libspx): plane edges alone, and two
    planes plus a caller-supplied quadratic term through the `extra` hook (how the ORB point edges would enter)"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "pose_check"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-o", str(exe),
                           os.path.join(root, "tests", "host", "pose_check.cpp")])
    out = subprocess.check_output([str(exe)], text=True).splitlines()
    a, b = out[0].split(), out[1].split()
    assert a[0] == "planes_only" and int(a[2]) == 0 and float(a[4]) < 1e-6
    assert b[0] == "with_prior" and int(b[2]) == 0 and float(b[4]) < 1e-6


def test_singular_start_is_left_alone():
    """With the camera axes exactly on the plane normals the azimuth of Plane3D is atan2(0, 0): the numeric Jacobian of
    the reference (restated) carries no information, every trial step is rejected and the pose comes back unchanged --
    documented behaviour of the parametrisation, not something to 'fix' on this side of the interface."""
    T_gt = pose([0.04, -0.08, 0.03], [0.2, -0.1, 0.3])
    e = api.plane_edges([0] * 6, ROOM[:6], measurements(T_gt, ROOM[:6]))
    T, outlier, chi2, bad = api.pose_optimize_planes(np.eye(4), e)
    assert np.array_equal(T, np.eye(4)) and bad > 0
    T2, _, _, bad2 = api.pose_optimize_planes(pose([1e-3, 2e-3, -1e-3], [0.01, 0, 0]), e)
    assert bad2 == 0 and np.abs(T2 - T_gt).max() < 1e-6


def test_degenerate_planes_are_refused():
    for bad_plane in ([0, 0, 0, 1.0], [np.nan, 0, 1, 1.0], [0, 0, 1, np.inf]):
        with pytest.raises(api.SpxError):
            api.pose_optimize_planes(np.eye(4), api.plane_edges([0], [bad_plane], [[0, 0, 1, 1.0]]))
        with pytest.raises(api.SpxError):
            api.pose_optimize_planes(np.eye(4), api.plane_edges([1], [[0, 0, 1, 1.0]], [bad_plane]))
    T = np.eye(4); T[0, 3] = np.nan
    with pytest.raises(api.SpxError):
        api.pose_optimize_planes(T, api.plane_edges([0], [[0, 0, 1, 1.0]], [[0, 0, 1, 1.0]]))
