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
