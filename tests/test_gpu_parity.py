"""-m gpu: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from sp_slam_b200 import api, scenes
from tests.parity import compare_frame

pytestmark = pytest.mark.gpu

FRAMES = [0, 80, 200, 240, 280, 600, 800, 880]


@pytest.fixture(scope="module")
def seq():
    P = scenes.poses(1000)
    return scenes.render(scenes.boxroom_rects(), P[FRAMES], scenes.TUM1)


@pytest.fixture(scope="module")
def ext():
    e = api.PlaneExtractor(debug=True, max_frames=8)
    yield e
    e.close()


@pytest.mark.parametrize("k", range(len(FRAMES)))
def test_boxroom_clean_frame(ext, seq, oracle_lib, k):
    d = seq[k]
    fp = ext.extract(d)
    orc = oracle_lib.Oracle().run(d)
    rep = compare_frame(ext, orc, d, fp)
    assert rep["normals_bit_exact"] and rep["labels_bit_exact"] and rep.get("models_bit_exact", True), rep


@pytest.mark.parametrize("k", [0, 2, 5])
def test_boxroom_noisy_frame(ext, seq, oracle_lib, k):
    d = scenes.add_noise(seq[k], FRAMES[k])
    fp = ext.extract(d)
    orc = oracle_lib.Oracle().run(d)
    rep = compare_frame(ext, orc, d, fp)
    assert rep["labels_bit_exact"], rep


def test_batch_equals_single(ext, seq):
    res = ext.extract_batch(seq)
    for k in range(len(FRAMES)):
        one = ext.extract(seq[k])
        b = res.frame(k)
        assert one.mnPlaneNum == b.mnPlaneNum and one.mnRealPlaneNum == b.mnRealPlaneNum
        assert np.array_equal(one.mvPlaneCoefficients.view(np.uint32), b.mvPlaneCoefficients.view(np.uint32))
        for p, q in zip(one.mvPlanePoints, b.mvPlanePoints):
            assert np.array_equal(p, q)
        for p, q in zip(one.mvBoundaryPoints, b.mvBoundaryPoints):
            assert np.array_equal(p, q)


def test_feed_oracle_normals_bit_exact_labels(ext, seq, oracle_lib):
    d = scenes.add_noise(seq[2], 200)
    orc = oracle_lib.Oracle().run(d)
    fp = ext.segment_from_normals(d, orc.normals())
    n = orc.n
    lab, nl = ext.labels_raw(0, n)
    lab_ref, nl_ref = orc.labels_raw()
    assert np.array_equal(lab, lab_ref.ravel()) and nl == nl_ref
    compare_frame(ext, orc, d, fp, check_stages=False)
