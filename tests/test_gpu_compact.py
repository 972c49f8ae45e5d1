"""-m gpu: compact results (ordered inlier index lists) and the C++ sequence adapter against the 16-byte cloud path.

The cloud path is itself compared with the oracle bit for bit (tests/test_gpu_parity.py); here every field the compact
path delivers -- after the host has rebuilt the real planes' clouds from the caller's depth image -- must equal it."""
import numpy as np
import pytest

from sp_slam_b200 import api, scenes

pytestmark = pytest.mark.gpu

FRAMES = [0, 80, 200, 240, 280, 600, 800, 880]


@pytest.fixture(scope="module")
def seq():
    P = scenes.poses(1000)
    return scenes.render(scenes.boxroom_rects(), P[FRAMES], scenes.TUM1)


def same_fields(a: api.FramePlanes, b: api.FramePlanes):
    assert a.mnRealPlaneNum == b.mnRealPlaneNum and a.mnPlaneNum == b.mnPlaneNum and a.flags == b.flags
    assert np.array_equal(a.mvPlaneCoefficients.view(np.uint32), b.mvPlaneCoefficients.view(np.uint32))
    assert np.array_equal(a.is_supposed, b.is_supposed) and np.array_equal(a.src, b.src)
    for i in range(a.mnPlaneNum):
        for x, y in ((a.mvPlanePoints[i], b.mvPlanePoints[i]), (a.mvBoundaryPoints[i], b.mvBoundaryPoints[i])):
            assert len(x) == len(y)
            for name in ("x", "y", "z"):
                assert np.array_equal(x[name].view(np.uint32), y[name].view(np.uint32)), (i, name)   # bits, so -0.0 != 0.0
            assert np.array_equal(x["rgba"], y["rgba"])


def test_compact_equals_clouds_single_frames(seq):
    ext = api.PlaneExtractor(max_frames=8)
    for k in range(len(FRAMES)):
        full = ext.extract_batch(seq[k:k + 1]).frame(0)
        cr = ext.extract_batch_compact(seq[k:k + 1])
        assert cr.index_width == 2 and (cr.cloud_width, cr.cloud_height, cr.cloud_dis) == (214, 160, 3)
        same_fields(cr.frame(0, seq[k], ext.cfg), full)
        # only the supposed planes' clouds travel as points
        h = cr.frames[0]
        pl = cr.planes[: h["n_planes"]]
        assert len(cr.points) == int(pl["n_points"][pl["is_supposed"] != 0].sum())
        assert len(cr.point_index) == int(pl["n_points"][pl["is_supposed"] == 0].sum())
    ext.close()


def test_compact_batch_groups_callback_and_device_path(seq):
    """A batch cut into several frame groups: per-group delivery through the callback, host offsets, the device path's
    spx_fetch_compact, and the mode guard of the fetch calls."""
    import torch
    n = 100
    depth = np.ascontiguousarray(np.concatenate([seq] * 13)[:n])
    host = torch.from_numpy(depth).pin_memory()
    ext = api.PlaneExtractor(max_frames=n)
    full = ext.extract_batch_ptr(host.data_ptr(), n, 480, 640, copy=True)
    seen, errors = [], []

    def on_group(f0, f1, view):
        # the frames of the group (and of every earlier group) are final in the view
        try:          # (an exception inside a ctypes callback would only be printed)
            for f in (f0, f1 - 1):
                same_fields(view.frame(f, depth[f], ext.cfg), full.frame(f))
        except Exception as e:
            errors.append(repr(e))
        seen.append((f0, f1))

    ext.set_group_callback(on_group)
    cr = ext.extract_batch_compact_ptr(host.data_ptr(), n, 480, 640, copy=True)
    ext.set_group_callback(None)
    assert not errors, errors
    assert len(seen) >= 2 and seen[0][0] == 0 and seen[-1][1] == n and all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
    for f in range(n):
        same_fields(cr.frame(f, depth[f], ext.cfg), full.frame(f))
    up, inplace, down = ext.transfer_bytes()
    assert down < 0.3 * (full.frames.nbytes + full.planes.nbytes + full.points.nbytes + full.boundary.nbytes)
    assert abs(down - cr.nbytes) <= 32 * 8 * 8
    # device-resident input
    dev = host.cuda()
    ext.set_result_mode(True)
    ext.extract_device(dev.data_ptr(), n, 480, 640)
    with pytest.raises(api.SpxError):
        ext.fetch()
    cd = ext.fetch_compact()
    for f in (0, 17, 50, 99):
        same_fields(cd.frame(f, depth[f], ext.cfg), full.frame(f))
    ext.set_result_mode(False)
    ext.extract_device(dev.data_ptr(), n, 480, 640)
    with pytest.raises(api.SpxError):
        ext.fetch_compact()
    same_fields(ext.fetch().frame(3), full.frame(3))
    ext.close()


def test_compact_u16_and_720p_wide_indices():
    """16-bit depth input; a 1280x720 image whose organized cloud (427 x 240 = 102 480 points) needs 32-bit indices."""
    P = scenes.poses(1000)
    d = scenes.render(scenes.boxroom_rects(), P[[200, 640]], scenes.TUM1)
    d16 = np.round(np.clip(scenes.add_noise(d[1], 640, "kinect"), 0, 13.0).astype(np.float64) * 5000.0).astype(np.uint16)[None]
    factor = float(np.float32(1.0) / np.float32(5000.0))
    ext = api.PlaneExtractor(max_frames=2)
    full = ext.extract_batch_u16(d16, factor).frame(0)
    cr = ext.extract_batch_u16_compact(d16, factor)
    same_fields(cr.frame(0, d16[0].astype(np.float32) * np.float32(factor), ext.cfg), full)
    ext.close()
    it = scenes.REALSENSE
    big = scenes.realsense_sequence(2, start=3)
    ext = api.PlaneExtractor(max_frames=2, max_rows=720, max_cols=1280, fx=it.fx, fy=it.fy, cx=it.cx, cy=it.cy,
                             max_x=float(it.width), max_y=float(it.height))
    full = ext.extract_batch(big)
    cr = ext.extract_batch_compact(big)
    assert cr.index_width == 4 and cr.cloud_width == 427
    for f in range(2):
        same_fields(cr.frame(f, big[f], ext.cfg), full.frame(f))
    ext.close()


def test_sequence_adapter_fills_every_frame_field(seq):
    """spx_host::SequencePlanes (C++ adapter, worker threads, streaming stores) against the C ABI's clouds: coefficients,
    32-byte points (data[3] = 1, padding zero), width / height conventions of the reference's clouds."""
    import torch
    n = 64
    depth = np.ascontiguousarray(np.concatenate([seq] * 8)[:n])
    host = torch.from_numpy(depth).pin_memory()
    ext = api.PlaneExtractor(max_frames=n)
    full = ext.extract_batch_ptr(host.data_ptr(), n, 480, 640, copy=True)
    ad = api.SequenceAdapter(api.default_config(max_frames=n), n_threads=4)
    hashes = set()
    for rep in range(3):     # pooled storage: later passes reuse the clouds of the first
        ms = ad.process_ptr(host.data_ptr(), n, 480, 640)
        assert ms > 0
        hashes.add(ad.hash())
    assert len(hashes) == 1
    n_pl, n_pt, n_bd, nbytes = ad.summary()
    assert n_pl == len(full.planes) and n_pt == len(full.points) and n_bd == len(full.boundary)
    for f in range(n):
        a, b = ad.frame(f), full.frame(f)
        same_fields_32(a, b)
    # the 16-byte-cloud transfer through the same adapter (group callback of spx_extract_batch): the same fields
    h0 = ad.hash()
    ad.process_clouds_ptr(host.data_ptr(), n, 480, 640)
    assert ad.hash() == h0
    # 16-bit input through the adapter
    d16 = np.round(np.clip(depth, 0, 13.0).astype(np.float64) * 5000.0).astype(np.uint16)
    h16 = torch.from_numpy(d16).pin_memory()
    factor = float(np.float32(1.0) / np.float32(5000.0))
    full16 = ext.extract_batch_u16(d16, factor)
    ad.process_u16_ptr(h16.data_ptr(), n, 480, 640, factor)
    for f in (0, 9, 63):
        same_fields_32(ad.frame(f), full16.frame(f))
    ad.close()
    ext.close()


def same_fields_32(a: api.FramePlanes, b: api.FramePlanes):
    assert a.mnRealPlaneNum == b.mnRealPlaneNum and a.mnPlaneNum == b.mnPlaneNum and a.flags == b.flags
    assert np.array_equal(a.mvPlaneCoefficients.view(np.uint32), b.mvPlaneCoefficients.view(np.uint32))
    for i in range(a.mnPlaneNum):
        for x, y in ((a.mvPlanePoints[i], b.mvPlanePoints[i]), (a.mvBoundaryPoints[i], b.mvBoundaryPoints[i])):
            assert len(x) == len(y)
            for name in ("x", "y", "z"):
                assert np.array_equal(x[name].view(np.uint32), y[name].view(np.uint32)), (i, name)
            assert np.array_equal(x["rgba"], y["rgba"])
            assert np.all(x["data_w"] == 1.0) and not x["pad"].any()
        (pw, ph), (bw, bh) = a.dims[i]
        n_p, n_b = len(b.mvPlanePoints[i]), len(b.mvBoundaryPoints[i])
        assert (pw, ph) == (n_p, 1)                         # ExtractIndices::filter / operator+= (src/Frame.cc:925,988)
        # real planes: `boundaryPoints->points = getContour()` leaves the cloud 0 x 0 (src/Frame.cc:930-932)
        assert (bw, bh) == ((n_b, 1) if b.is_supposed[i] else (0, 0))
