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
