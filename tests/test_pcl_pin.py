"""Pinning the oracle to REAL PCL: consumes tests/golden/pcl_*.npz when present.

Those files are produced by tools/pcl_pin (a C++ program that calls pcl::IntegralImageNormalEstimation,
pcl::OrganizedMultiPlaneSegmentation::segment / segmentAndRefine and pcl::SACSegmentation exactly as
/root/reference/src/Frame.cc:878-905,939-995 does and dumps every intermediate).  PCL cannot be built in this
repository's container, so no such file is committed yet and the first test skips; the day someone with PCL 1.8 runs the
recipe in tools/pcl_pin/pcl_pin.cpp, parity of the oracle (and through tests/test_gpu_parity.py of the CUDA path) is
pinned or the disagreement is localised stage by stage -- `orc_config.alt` (ORC_ALT_*) flips the recalled PCL details
one at a time to bisect it.

The second test keeps the comparer honest meanwhile: a stand-in dump written from the oracle itself in the same format
must pass, and must fail when a single value of any stage is changed.
"""
import glob
import os

import numpy as np
import pytest

from oracle import pyoracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_f32(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))))


def run_oracle(g, alt=0):
    fx, fy, cx, cy = (float(v) for v in g["intrinsics"])
    rows, cols = g["depth"].shape
    return pyoracle.Oracle(fx=fx, fy=fy, cx=cx, cy=cy, max_x=float(cols), max_y=float(rows), alt=alt).run(g["depth"])


def lists(g, stem):
    val, off = g[stem + "_val"], g[stem + "_off"]
    return [val[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def compare_with_pcl(g, orc):
    """Stage-by-stage differences between the oracle run `orc` and the PCL dump `g`; an empty dict = pinned."""
    diff = {}
    w, h = (int(v) for v in g["dims"])
    if (w, h) != (orc.width, orc.height):
        return {"dims": ((w, h), (orc.width, orc.height))}
    if not same_f32(g["cloud"], orc.cloud()):
        diff["cloud"] = int((bits(g["cloud"]) != bits(orc.cloud())).sum())
    n_ref, n_orc = g["normals"][:3], orc.normals()
    if not np.array_equal(np.isnan(n_ref[0]), np.isnan(n_orc[0])):
        diff["normal_validity"] = int((np.isnan(n_ref[0]) != np.isnan(n_orc[0])).sum())
    elif not same_f32(n_ref, n_orc):
        ok = ~np.isnan(n_ref[0])
        a, b = n_ref[:, ok].astype(np.float64), n_orc[:, ok].astype(np.float64)
        ang = np.arctan2(np.linalg.norm(np.cross(a.T, b.T), axis=1), np.sum(a * b, axis=0))
        diff["normals"] = {"pixels": int((bits(n_ref) != bits(n_orc)).any(axis=0).sum()), "max_angle_rad": float(ang.max())}
    if not np.all(np.isnan(g["normals"][3])):
        diff["curvature_not_nan"] = int((~np.isnan(g["normals"][3])).sum())        # AVERAGE_3D_GRADIENT leaves curvature NaN
    lab, n_lists = orc.labels_raw()
    if not np.array_equal(g["seg_labels"].astype(np.uint32), lab.ravel()):
        diff["seg_labels"] = int((g["seg_labels"].astype(np.uint32) != lab.ravel()).sum())
    if int(g["seg_n_label_lists"][0]) != n_lists:
        diff["n_label_lists"] = (int(g["seg_n_label_lists"][0]), n_lists)
    models = orc.models()
    if len(g["seg_coef"]) != len(models) or len(g["ref_coef"]) != len(models):
        diff["model_count"] = (len(g["seg_coef"]), len(g["ref_coef"]), len(models))
        return diff
    seg_inl, ref_inl = lists(g, "seg_inliers"), lists(g, "ref_inliers")
    for i, m in enumerate(models):
        for key, ref in (("coef", g["seg_coef"][i]), ("coef_after_refine", g["ref_coef"][i]), ("centroid", g["seg_centroids"][i][:3]),
                         ("cov", g["seg_cov"][i].reshape(3, 3))):
            mine = m["coef"] if key.startswith("coef") else m[key]
            if not same_f32(ref, mine):
                diff.setdefault("models", {})[f"{i}.{key}"] = (np.asarray(ref).ravel().tolist(), np.asarray(mine).ravel().tolist())
        if not np.array_equal(seg_inl[i], m["inliers"][: m["n_segment"]]):
            diff.setdefault("models", {})[f"{i}.segment_inliers"] = (len(seg_inl[i]), m["n_segment"])
        if not np.array_equal(ref_inl[i], m["inliers"]):
            diff.setdefault("models", {})[f"{i}.refined_inliers"] = (len(ref_inl[i]), len(m["inliers"]))
        c0, c1 = int(g["contour_off"][i]), int(g["contour_off"][i + 1])
        cl = orc.cloud()
        mine = cl[:, m["contour"]].T if len(m["contour"]) else np.zeros((0, 3), np.float32)
        if not same_f32(g["contour_pts"][c0:c1], mine):
            diff.setdefault("models", {})[f"{i}.contour"] = (c1 - c0, len(m["contour"]))
    if not np.array_equal(g["ref_labels"].astype(np.uint32), orc.labels_refined().ravel()):
        diff["ref_labels"] = int((g["ref_labels"].astype(np.uint32) != orc.labels_refined().ravel()).sum())
    # RANSAC line fits: the oracle logs one record per SACSegmentation::segment call of the planes it keeps; PCL's dump has
    # every region with a contour of at least 50 points, keyed by (model, round)
    pcl_lines = {(int(r[0]), int(r[1])): (int(r[2]), int(r[3]), g["line_coef"][k]) for k, r in enumerate(g["line_rec"])}
    planes = orc.planes()
    for rec in orc.line_recs():
        key = (planes[rec["plane"]]["src"], rec["round"])
        if key not in pcl_lines:
            diff.setdefault("lines", {})[str(key)] = "missing in the PCL dump"
            continue
        n_pts, n_inl, coef = pcl_lines[key]
        if (n_pts, n_inl) != (rec["n_points"], rec["n_inliers"]) or not same_f32(coef, rec["coef"]):
            diff.setdefault("lines", {})[str(key)] = ((n_pts, n_inl, coef.tolist()), (rec["n_points"], rec["n_inliers"], rec["coef"].tolist()))
    return diff


def standin_dump(depth, intr, alt=0):
    """What tools/pcl_pin would write if PCL behaved exactly like the oracle (with the alternative readings `alt`)."""
    rows, cols = depth.shape
    orc = pyoracle.Oracle(fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3], max_x=float(cols), max_y=float(rows), alt=alt).run(depth)
    models = orc.models()
    lab, n_lists = orc.labels_raw()
    cl = orc.cloud()

    def flat(ls):
        off = np.zeros(len(ls) + 1, np.int32)
        off[1:] = np.cumsum([len(x) for x in ls])
        return (np.concatenate(ls).astype(np.int32) if ls else np.zeros(0, np.int32)), off

    g = {"depth": depth, "intrinsics": np.array(intr, np.float32), "dims": np.array([orc.width, orc.height], np.int32), "cloud": cl,
         "normals": np.concatenate([orc.normals(), np.full((1, orc.n), np.nan, np.float32)]),
         "seg_labels": lab.ravel().astype(np.int32), "seg_n_label_lists": np.array([n_lists], np.int32),
         "seg_coef": np.array([m["coef"] for m in models], np.float32).reshape(-1, 4),
         "ref_coef": np.array([m["coef"] for m in models], np.float32).reshape(-1, 4),
         "seg_centroids": np.array([list(m["centroid"]) + [1.0] for m in models], np.float32).reshape(-1, 4),
         "seg_cov": np.array([m["cov"].ravel() for m in models], np.float32).reshape(-1, 9),
         "ref_labels": orc.labels_refined().ravel().astype(np.int32)}
    g["seg_inliers_val"], g["seg_inliers_off"] = flat([m["inliers"][: m["n_segment"]] for m in models])
    g["ref_inliers_val"], g["ref_inliers_off"] = flat([m["inliers"] for m in models])
    con = [cl[:, m["contour"]].T.astype(np.float32) for m in models]
    g["contour_pts"] = np.concatenate(con) if con else np.zeros((0, 3), np.float32)
    g["contour_off"] = np.concatenate([[0], np.cumsum([len(c) for c in con])]).astype(np.int32)
    planes = orc.planes()
    recs = orc.line_recs()
    g["line_rec"] = np.array([[planes[r["plane"]]["src"], r["round"], r["n_points"], r["n_inliers"]] for r in recs], np.int32).reshape(-1, 4)
    g["line_coef"] = np.array([r["coef"] for r in recs], np.float32).reshape(-1, 6)
    return g


def test_oracle_against_real_pcl_dumps():
    files = sorted(glob.glob(os.path.join(GOLDEN, "pcl_*.npz")))
    if not files:
        pytest.skip("no tests/golden/pcl_*.npz: PCL cannot be built in this container; run the recipe in tools/pcl_pin/pcl_pin.cpp "
                    "on a machine with PCL 1.8 to pin the oracle (parity unpinned until then)")
    for path in files:
        g = np.load(path)
        diff = compare_with_pcl(g, run_oracle(g))
        if diff:
            # which of the recalled details explains it?
            hints = {name: not compare_with_pcl(g, run_oracle(g, alt))
                     for name, alt in (("ALT_VP_RESET", 1), ("ALT_CHAMFER_NO_WRAP", 2), ("ALT_REFINE_NO_WRAP", 4), ("ALT_SAMPLE_GOOD_OR", 8), ("ALT_RNG_MASK", 16))}
            raise AssertionError(f"{os.path.basename(path)}: the oracle disagrees with PCL: {diff}; single switches that fix it: {hints}")


def test_comparer_on_a_standin_dump():
    from sp_slam_b200 import scenes
    depth = scenes.render(scenes.boxroom_rects(), scenes.poses(1000)[[200]], scenes.TUM1)[0]
    it = scenes.TUM1
    intr = (it.fx, it.fy, it.cx, it.cy)
    g = standin_dump(depth, intr)
    orc = run_oracle(g)
    assert len(orc.models()) >= 2 and len(orc.line_recs()) >= 1
    assert compare_with_pcl(g, orc) == {}
    # every stage is really looked at: one changed value per stage must be reported
    for key, poke in (("normals", lambda a: a.__setitem__((0, 5000 + int(np.flatnonzero(~np.isnan(a[0, 5000:]))[0])), 0.123)),
                      ("seg_labels", lambda a: a.__setitem__(17000, a[17000] + 1)),
                      ("seg_cov", lambda a: a.__setitem__((0, 4), a[0, 4] * 1.0000002)),
                      ("ref_inliers_val", lambda a: a.__setitem__(-1, a[-1] ^ 1)),
                      ("contour_pts", lambda a: a.__setitem__((3, 2), a[3, 2] + 1e-3)),
                      ("line_coef", lambda a: a.__setitem__((0, 3), -a[0, 3]))):
        h = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in g.items()}
        poke(h[key])
        assert compare_with_pcl(h, orc) != {}, key
    # the alternative readings are live switches: a dump written under another reading is told apart from the default,
    # and matches again when the oracle runs with the same switch
    other = scenes.render(scenes.boxroom_rects(), scenes.poses(1000)[[880]], scenes.TUM1)[0]
    told_apart = set()
    for alt in (pyoracle.ALT_VP_RESET, pyoracle.ALT_CHAMFER_NO_WRAP, pyoracle.ALT_REFINE_NO_WRAP, pyoracle.ALT_SAMPLE_GOOD_OR, pyoracle.ALT_RNG_MASK):
        ga = standin_dump(other, intr, alt)
        assert compare_with_pcl(ga, run_oracle(ga, alt)) == {}
        if compare_with_pcl(ga, run_oracle(ga)) != {}:
            told_apart.add(alt)
    # on this frame the viewpoint accumulation, isSampleGood and the RNG reading each change a result; the two row-wrap
    # readings only matter next to the image border (tests/test_oracle_kat.py pins the chamfer one on a crafted mask)
    assert told_apart >= {pyoracle.ALT_VP_RESET, pyoracle.ALT_SAMPLE_GOOD_OR, pyoracle.ALT_RNG_MASK}
