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


def test_cpp_host_adapter_fills_frame_fields(ext, seq, tmp_path):
    """The C++ adapter (sp_slam_b200/host/FramePlanes.h, the reference's member names) against the C ABI."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "adapter_check"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", str(exe), os.path.join(root, "tests", "host", "adapter_check.cpp"),
                           "-L" + os.path.join(root, "sp_slam_b200"), "-lspx", "-Wl,-rpath," + os.path.join(root, "sp_slam_b200")])
    d = seq[2]
    raw = tmp_path / "depth.bin"
    d.tofile(raw)
    out = subprocess.check_output([str(exe), str(raw), "480", "640"], text=True).splitlines()
    fp = ext.extract(d)
    assert out[0] == f"real {fp.mnRealPlaneNum}" and out[1] == f"all {fp.mnPlaneNum}"
    for i, line in enumerate(out[2:]):
        t = line.split()
        assert np.array_equal(np.array(t[2:6], np.float32).view(np.uint32), fp.mvPlaneCoefficients[i].view(np.uint32))
        assert int(t[6]) == len(fp.mvPlanePoints[i]) and int(t[7]) == len(fp.mvBoundaryPoints[i])
        p = fp.mvPlanePoints[i]
        sx = float(np.sum(p["x"].astype(np.float64) + 2.0 * p["y"].astype(np.float64) + 3.0 * p["z"].astype(np.float64)))
        assert abs(float(t[8]) - sx) <= 1e-9 * max(1.0, abs(sx))


def test_frame_groups_host_and_device_paths_agree():
    """A batch large enough to be cut into several frame groups (internal streams, per-group compaction and
    download on the host path; batch-wide compaction on the device path) returns what the one-frame call returns."""
    import torch
    n = 100
    d = scenes.boxroom_sequence(n, start=170)
    ext = api.PlaneExtractor(max_frames=n, n_streams=4)
    host = ext.extract_batch(d)
    dev = torch.from_numpy(d).cuda()
    ext.extract_device(dev.data_ptr(), n, 480, 640)
    devres = ext.fetch()
    one = api.PlaneExtractor()
    assert len(host) == len(devres) == n
    assert int(host.frames["n_planes"].sum()) == len(host.planes) == len(devres.planes)
    for k in range(n):
        a, b = host.frame(k), devres.frame(k)
        ref = one.extract(d[k]) if k % 9 == 0 else None
        for other in (b, ref):
            if other is None:
                continue
            assert a.mnPlaneNum == other.mnPlaneNum and a.mnRealPlaneNum == other.mnRealPlaneNum
            assert np.array_equal(a.mvPlaneCoefficients.view(np.uint32), other.mvPlaneCoefficients.view(np.uint32))
            for p, q in zip(a.mvPlanePoints, other.mvPlanePoints):
                assert np.array_equal(p, q)
            for p, q in zip(a.mvBoundaryPoints, other.mvBoundaryPoints):
                assert np.array_equal(p, q)
    ext.close()
    one.close()
