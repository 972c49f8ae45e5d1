"""Stage-by-stage comparison of the CUDA path (through the C ABI) with the CPU oracle on one frame."""
from __future__ import annotations

import numpy as np


def angle_between(n1, n2):
    a, b = np.asarray(n1, np.float64), np.asarray(n2, np.float64)
    return float(np.arctan2(np.linalg.norm(np.cross(a, b)), np.dot(a, b)))   # well conditioned near 0


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_f32(a, b):
    """bit-identical, NaN == NaN regardless of payload"""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))))


def ext_flags(frame_planes):
    return int(getattr(frame_planes, "flags", 0) or 0) if frame_planes is not None else 0


def compare_frame(ext, orc, depth, frame_planes=None, tol_angle=1e-4, tol_d=1e-4, check_stages=True, frame=0, expect_sat_exact=True):
    """ext: PlaneExtractor(debug=True) that has just processed `depth` as frame `frame`; orc: Oracle already run on it.
    Returns a dict of findings; raises AssertionError on a parity violation."""
    rows, cols = depth.shape
    w, h = ext.cloud_dims(rows, cols)
    assert (w, h) == (orc.width, orc.height)
    n = w * h
    rep = {}
    # the claim the tile- / strip-local integral images rest on: every fp64 partial sum of this frame was exact in the
    # oracle's whole-image integral images (error-free-transformation check), so the summation order cannot matter
    rep["sat_exact"] = bool(orc.sat_exact())
    if check_stages and expect_sat_exact:
        assert rep["sat_exact"], "the oracle's integral-image sums rounded on this frame: bit-identity with PCL is not defined"
        assert not (ext_flags(frame_planes) & 4), "SPX_FRAME_SAT_UNPROVEN set on a frame whose sums are exact and well inside the bound"
    if check_stages:
        assert same_f32(ext.cloud(frame, n), orc.cloud()), "organized cloud differs"
        d_gpu = ext.distance_map(frame, n)
        d_ref = np.minimum(orc.distance_map().ravel(), np.float32(10.0))
        assert same_f32(d_gpu, d_ref), f"distance map differs at {np.flatnonzero(bits(d_gpu) != bits(d_ref))[:10]}"
        nrm_gpu, pd_gpu = ext.normals(frame, n)
        nrm_ref = orc.normals()
        nan_g, nan_r = np.isnan(nrm_gpu[0]), np.isnan(nrm_ref[0])
        assert np.array_equal(nan_g, nan_r), "normal validity differs"
        rep["normals_bit_exact"] = same_f32(nrm_gpu, nrm_ref)
        if not rep["normals_bit_exact"]:
            ok = ~nan_r
            a, b = nrm_gpu[:, ok].astype(np.float64), nrm_ref[:, ok].astype(np.float64)
            rep["normals_max_angle"] = float(np.arctan2(np.linalg.norm(np.cross(a.T, b.T), axis=1), np.sum(a * b, axis=0)).max())
            assert rep["normals_max_angle"] < 1e-4, rep
        else:
            assert same_f32(pd_gpu, orc.plane_d()), "plane_d differs"
        lab_gpu, nl_gpu = ext.labels_raw(frame, n)
        lab_ref, nl_ref = orc.labels_raw()
        rep["labels_bit_exact"] = bool(np.array_equal(lab_gpu, lab_ref.ravel())) and nl_gpu == nl_ref
        if rep["normals_bit_exact"]:
            assert rep["labels_bit_exact"], "raw CCL labels differ although the normals are bit-identical"
    # models
    mg, mr = ext.models(frame), orc.models()
    assert len(mg) == len(mr), f"model count {len(mg)} != {len(mr)}"
    ids = ext.plane_ids(frame, n)
    lab_refined = orc.labels_refined().ravel()
    for i, (a, b) in enumerate(zip(mg, mr)):
        assert a["label"] == b["label"] and a["n_segment"] == b["n_segment"], (i, a["label"], b["label"])
        ang = angle_between(a["coef"][:3], b["coef"][:3])
        assert ang < tol_angle, f"model {i}: normal differs by {ang} rad"
        assert abs(float(a["coef"][3]) - float(b["coef"][3])) < tol_d, f"model {i}: offset"
        rep.setdefault("models_bit_exact", True)
        if not (same_f32(a["coef"], b["coef"]) and same_f32(a["cov"], b["cov"]) and same_f32(a["centroid"], b["centroid"])):
            rep["models_bit_exact"] = False
        # per-plane IoU of the refined label maps
        g = ids == i
        r = lab_refined == b["label"]
        iou = (g & r).sum() / max((g | r).sum(), 1)
        assert iou >= 0.995, f"model {i}: IoU {iou}"
        rep.setdefault("min_iou", 1.0)
        rep["min_iou"] = min(rep["min_iou"], float(iou))
        if rep.get("models_bit_exact", False):
            assert np.array_equal(a["inliers"], b["inliers"]), f"model {i}: inlier list (order) differs"
            assert np.array_equal(a["contour"], b["contour"]), f"model {i}: contour differs"
    # final Frame fields
    if frame_planes is not None:
        pr = orc.planes()
        assert frame_planes.mnRealPlaneNum == orc.n_real, (frame_planes.mnRealPlaneNum, orc.n_real)
        assert frame_planes.mnPlaneNum == orc.n_planes, (frame_planes.mnPlaneNum, orc.n_planes)
        exact = rep.get("models_bit_exact", True)
        for i, b in enumerate(pr):
            cg = frame_planes.mvPlaneCoefficients[i]
            assert cg[3] >= 0
            ang = angle_between(cg[:3], b["coef"][:3])
            assert ang < tol_angle and abs(float(cg[3]) - float(b["coef"][3])) < tol_d, (i, cg, b["coef"])
            assert int(frame_planes.src[i]) == b["src"]
            assert len(frame_planes.mvPlanePoints[i]) == len(b["points"]), (i, len(frame_planes.mvPlanePoints[i]), len(b["points"]))
            assert len(frame_planes.mvBoundaryPoints[i]) == len(b["boundary"])
            if exact:
                assert same_f32(cg, b["coef"]), (i, cg, b["coef"])
                for fld in ("x", "y", "z"):
                    assert same_f32(frame_planes.mvPlanePoints[i][fld], b["points"][fld]), (i, fld, "points")
                    assert same_f32(frame_planes.mvBoundaryPoints[i][fld], b["boundary"][fld]), (i, fld, "boundary")
                assert np.array_equal(frame_planes.mvPlanePoints[i]["rgba"], b["points"]["rgba"])
                assert np.array_equal(frame_planes.mvBoundaryPoints[i]["rgba"], b["boundary"]["rgba"])
        if exact:
            lg, lr = ext.lines(frame), orc.line_recs()
            assert len(lg) == len(lr), (len(lg), len(lr))
            for a, b in zip(lg, lr):
                for k in ("plane", "round", "n_points", "iterations", "n_inliers", "in_range", "is_border", "emitted"):
                    assert int(a[k]) == int(b[k]), (k, a, b)
                assert same_f32(a["coef"], b["coef"]), (a["coef"], b["coef"])
        rep["n_real"], rep["n_planes"] = orc.n_real, orc.n_planes
    return rep
