"""Resident 1000-frame step under different group schedules (development aid):
    python tools/pipe_sweep.py frames streams [timeline]      (reads the SPX_* tuning knobs from the environment)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sp_slam_b200 import api, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
streams = int(sys.argv[2]) if len(sys.argv) > 2 else 8
timeline = len(sys.argv) > 3 and sys.argv[3] == "1"
d = scenes.boxroom_sequence(n)
dev = torch.from_numpy(d).cuda()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ext = api.PlaneExtractor(max_frames=n, n_streams=streams)
ext.set_stream(stream.cuda_stream)
for _ in range(3):
    ext.extract_device(dev.data_ptr(), n, 480, 640)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    reps = 5
    for _ in range(reps):
        ext.extract_device(dev.data_ptr(), n, 480, 640)
    e1.record(stream)
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / reps)
knobs = {k: v for k, v in os.environ.items() if k.startswith("SPX_")}
print("streams %d %s: step %.3f ms -> %.0f frames/s" % (streams, knobs, best, n / best * 1e3), flush=True)
if timeline:
    ext.set_profile(True)
    ext.extract_device(dev.data_ptr(), n, 480, 640)
    torch.cuda.synchronize()
    for row in ext.kernel_timeline():
        print(row)
