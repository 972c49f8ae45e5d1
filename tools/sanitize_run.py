"""Small end-to-end workload for compute-sanitizer: host path (groups, u16), device path, taps, 720p frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sp_slam_b200 import api, scenes

d = scenes.boxroom_sequence(66, start=190)
d[5, 100:140, 200:260] = np.nan
ext = api.PlaneExtractor(max_frames=66, n_streams=2, debug=True)
r = ext.extract_batch(d)
u16 = np.round(np.clip(np.nan_to_num(d), 0, 13.0).astype(np.float64) * 5000.0).astype(np.uint16)
r2 = ext.extract_batch_u16(u16)
dev = torch.from_numpy(d).cuda()
ext.extract_device(dev.data_ptr(), 66, 480, 640)
r3 = ext.fetch()
ext.models(3); ext.lines(3); ext.labels_raw(3, 214 * 160)
ext.close()
it = scenes.REALSENSE
e2 = api.PlaneExtractor(max_rows=720, max_cols=1280, fx=it.fx, fy=it.fy, cx=it.cx, cy=it.cy, max_x=1280.0, max_y=720.0)
f = e2.extract(scenes.add_noise(scenes.realsense_sequence(1, start=100)[0], 100, "realsense"))
e2.close()
os.environ["SPX_REFINE_FAST_MAX"] = "0"
e3 = api.PlaneExtractor(max_frames=8)
e3.extract_batch(d[:8])
e3.close()
print("planes", len(r.planes), len(r2.planes), len(r3.planes), f.mnPlaneNum)
# sparse upload (page-locked input: k_border_fetch, strided copies, early downloads), float and 16-bit, several groups
os.environ.pop("SPX_REFINE_FAST_MAX", None)
dd = np.ascontiguousarray(np.nan_to_num(d))
hp = torch.from_numpy(dd).pin_memory()
hp16 = torch.from_numpy(u16).pin_memory()
e4 = api.PlaneExtractor(max_frames=66, n_streams=3)
s1 = e4.extract_batch_ptr(hp.data_ptr(), 66, 480, 640, copy=True)
s2 = e4.extract_batch_u16_ptr(hp16.data_ptr(), 66, 480, 640, float(np.float32(1.0) / np.float32(5000.0)), copy=True)
print("sparse", len(s1.planes), len(s2.planes), e4.transfer_bytes())
# voxel grid (host clouds and device results), map upload / associate / boundary updates
fr = s1.frame(10)
outs = e4.voxel_grid(fr.mvPlanePoints + fr.mvBoundaryPoints + [np.empty(0, api.POINT_DTYPE)], 0.03)
e4.extract_device(dev.data_ptr(), 66, 480, 640)
e4.voxel_downsample_results(0.05, 1)
r5 = e4.fetch()
pm = api.PlaneMap(e4)
pm.upload(fr.mvPlaneCoefficients, fr.mvBoundaryPoints)
a = pm.associate(s1.frame(11).mvPlaneCoefficients)
pm.update_boundary(0, np.eye(4), np.concatenate(fr.mvBoundaryPoints))
e4.extract_device(dev.data_ptr(), 66, 480, 640)
f12 = r3.frame(12)
pm.update_boundary_from_result(1 % fr.mnPlaneNum, np.eye(4), 12, 0, len(f12.mvBoundaryPoints[0]))
a2 = pm.associate(s1.frame(12).mvPlaneCoefficients)
pm.close(); e4.close()
print("next rows", [len(o) for o in outs][:4], len(r5.boundary), a[0], a2[0])
# gathered upload route (host threads stage the samples; second tensor map; border windows fetch every row), every group gathered
# and the mixed mode, compact results; the ROI case whose cloud width is not a multiple of four; k_refine2's data-flow hand-over
e5 = api.PlaneExtractor(max_frames=66, n_streams=3)
e5.set_upload_mode(3)
e5.set_gather_threads(3)
g1 = e5.extract_batch_compact_ptr(hp.data_ptr(), 66, 480, 640)
g2 = e5.extract_batch_ptr(hp.data_ptr(), 66, 480, 640, copy=True)
os.environ["SPX_GATHER_MIN_MB"] = "0"
e6 = api.PlaneExtractor(max_frames=66, n_streams=3)
os.environ.pop("SPX_GATHER_MIN_MB", None)
g3 = e6.extract_batch_compact_ptr(hp.data_ptr(), 66, 480, 640)
print("gathered", len(g1.planes), len(g2.planes), len(g3.planes), e5.transfer_bytes()[0], e6.transfer_bytes()[0])
e5.close(); e6.close()
pad = np.zeros((4, 482, 700), np.float32)
pad[:, :401, :500] = dd[:4, :401, :500]
api.host_register(pad)
e7 = api.PlaneExtractor(max_frames=4, max_rows=401, max_cols=500, max_x=500.0, max_y=401.0)
e7.set_upload_mode(3)
g4 = e7.extract_batch(pad[:, :401, :500])
print("gathered roi", len(g4.planes), e7.transfer_bytes()[0] == 4 * 134 * 168 * 4)
e7.close()
api.host_unregister(pad)
