"""The committed golden fixtures (tests/golden, made by tools/make_golden.py from the oracle): on CPU they pin the
oracle + renderer; on the GPU (-m gpu) the CUDA path has to reproduce them through the C ABI."""
import hashlib

import numpy as np
import pytest

from tests.golden_util import CASES, load


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(oracle_lib, name):
    g, depth = load(name)
    assert sha(depth) == str(g["depth_sha"]), "input synthesis drifted"
    o = oracle_lib.Oracle().run(depth)
    assert (o.n_real, o.n_planes) == (int(g["n_real"]), int(g["n_planes"]))
    pls = o.planes()
    assert np.array_equal(np.stack([p["coef"] for p in pls]).view(np.uint32), g["coef"].view(np.uint32))
    assert [sha(p["points"]) for p in pls] == list(g["points_sha"])
    assert [sha(p["boundary"]) for p in pls] == list(g["boundary_sha"])
    assert sha(o.labels_raw()[0]) == str(g["labels_raw_sha"]) and o.labels_raw()[1] == int(g["n_label_lists"])
    assert sha(o.labels_refined()) == str(g["labels_refined_sha"])
    assert sha(o.normals()) == str(g["normals_sha"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_golden(name):
    from sp_slam_b200 import api
    g, depth = load(name)
    ext = api.PlaneExtractor(debug=True)
    fp = ext.extract(depth)
    n = 214 * 160
    assert (fp.mnRealPlaneNum, fp.mnPlaneNum) == (int(g["n_real"]), int(g["n_planes"]))
    assert np.array_equal(fp.mvPlaneCoefficients.view(np.uint32), g["coef"].view(np.uint32))
    assert [sha(p) for p in fp.mvPlanePoints] == list(g["points_sha"])
    assert [sha(p) for p in fp.mvBoundaryPoints] == list(g["boundary_sha"])
    assert np.array_equal(fp.src, g["src"])
    lab, n_lists = ext.labels_raw(0, n)
    assert sha(lab) == str(g["labels_raw_sha"]) and n_lists == int(g["n_label_lists"])
    nrm, _ = ext.normals(0, n)
    assert sha(nrm) == str(g["normals_sha"])
    assert sha(ext.distance_map(0, n).reshape(160, 214)) == str(g["distance_sha"])
    lines = ext.lines(0)
    rec = np.array([[int(l[k]) for k in ("plane", "round", "n_points", "iterations", "n_inliers", "in_range",
                                          "is_border", "emitted")] for l in lines], np.int32).reshape(-1, 8)
    assert np.array_equal(rec, g["line_recs"])
    assert np.array_equal(np.array([l["coef"] for l in lines], np.float32).reshape(-1, 6).view(np.uint32),
                          g["line_coef"].view(np.uint32))
    ext.close()


# ---- one fixture with inputs and oracle outputs (tools/make_golden.py: next_rows_fixture) ----
def _next_rows():
    import os
    from tests.golden_util import GOLDEN_DIR
    return np.load(os.path.join(GOLDEN_DIR, "next_rows.npz"))


def test_oracle_reproduces_next_rows_golden(oracle_lib):
    g = _next_rows()
    for leaf in (0.01, 0.05):
        out, idx = oracle_lib.voxel_grid(g["vox_in"], leaf)
        assert np.array_equal(out, g[f"vox_out_{leaf}"]) and np.array_equal(idx, g[f"vox_idx_{leaf}"])
    off = g["map_off"]
    bnds = [g["map_bnd"][off[j]:off[j + 1]] for j in range(len(g["map_w"]))]
    got = oracle_lib.associate_planes(g["plane_w"], g["map_w"], bnds, n_seen=int(g["n_seen"]))
    for a, b in zip(got, (g["assoc"], g["vertical"], g["parallel"], g["assoc_dist"])):
        assert np.array_equal(a, b)
    assert np.array_equal(oracle_lib.transform_cloud(bnds[int(g["xform_src"])], g["xform"]), g["xform_out"])


@pytest.mark.gpu
def test_cuda_reproduces_next_rows_golden():
    from sp_slam_b200 import api
    g = _next_rows()
    ext = api.PlaneExtractor()
    for leaf in (0.01, 0.05):
        out = ext.voxel_grid([g["vox_in"]], leaf)[0]
        ref = g[f"vox_out_{leaf}"]
        assert len(out) == len(ref) and np.array_equal(out["rgba"], ref["rgba"])
        for ax in "xyz":      # members of a voxel are summed in input order here, in std::sort's order there: last-ulp agreement
            assert np.allclose(out[ax], ref[ax], rtol=0, atol=64 * 2.0 ** -23 * 3.0)
            assert (out[ax] == ref[ax]).mean() > 0.8
    off = g["map_off"]
    bnds = [g["map_bnd"][off[j]:off[j + 1]] for j in range(len(g["map_w"]))]
    pm = api.PlaneMap(ext)
    pm.upload(g["map_w"], bnds, int(g["n_seen"]))
    got = pm.associate(g["plane_w"])
    for a, b in zip(got, (g["assoc"], g["vertical"], g["parallel"], g["assoc_dist"])):
        assert np.array_equal(a, b)
    src = int(g["xform_src"])
    pm.update_boundary(0, g["xform"], bnds[src])
    assert np.array_equal(pm.boundary(0), g["xform_out"])
    pm.close(); ext.close()
