"""The host half of the gathered upload (spx_host_gather_samples: GatherPool of spx_api.cu) -- runs without a GPU."""
import numpy as np

from sp_slam_b200 import api


def test_gather_equals_strided_view():
    """out[f, m, n] = depth[f, m * dis, n * dis] (src/Frame.cc:857-872) whatever the group / thread split; row tails zeroed"""
    rng = np.random.default_rng(3)
    d = rng.random((37, 96, 160), dtype=np.float32)
    want = d[:, ::3, ::3]
    for groups, threads in ((1, 1), (4, 3), (8, 8), (37, 5), (50, 2)):
        got = api.gather_samples(d, 3, groups, threads)
        assert got.shape == (37, 32, 56) and np.array_equal(got[:, :, :54], want) and not got[:, :, 54:].any(), (groups, threads)


def test_gather_pitched_roi_and_odd_sizes():
    rng = np.random.default_rng(4)
    big = rng.random((9, 123, 331), dtype=np.float32)
    roi = big[:, 3:120, 7:326]          # 117 x 319, pitched, frame stride != rows * pitch
    for dis in (1, 2, 3, 5, 7):
        w, h = -(-319 // dis), -(-117 // dis)
        got = api.gather_samples(roi, dis, 3, 4)
        assert got.shape == (9, h, (w + 3) & ~3)
        assert np.array_equal(got[:, :, :w], roi[:, ::dis, ::dis]) and not got[:, :, w:].any(), dis
    one = api.gather_samples(roi[4], 5, 1, 2, row_floats=70)
    assert one.shape == (1, 24, 70) and np.array_equal(one[0, :, :64], roi[4, ::5, ::5])


def test_gather_rejects_bad_arguments():
    d = np.zeros((2, 8, 8), np.float32)
    L = api.lib()
    out = np.zeros((2, 3, 4), np.float32)
    assert L.spx_host_gather_samples(d.ctypes.data, 2, 8, 8, 32, 256, 3, 1, 0, out.ctypes.data, 4) != api.SPX_OK      # no threads
    assert L.spx_host_gather_samples(d.ctypes.data, 2, 8, 8, 32, 256, 3, 1, 1, out.ctypes.data, 2) != api.SPX_OK      # rows too short
    assert L.spx_host_gather_samples(None, 2, 8, 8, 32, 256, 3, 1, 1, out.ctypes.data, 4) != api.SPX_OK
    assert L.spx_host_gather_samples(d.ctypes.data, 2, 8, 8, 32, 256, 3, 1, 1, out.ctypes.data, 4) == api.SPX_OK
    assert np.array_equal(out[:, :, :3], d[:, ::3, ::3])
