"""Per-kernel CUDA-event times of a resident 1000-frame step run as ONE frame group, and the wall time of the step with the
default frame groups (development aid): python tools/ktimes.py [frames]   (reads the SPX_* tuning knobs from the environment)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sp_slam_b200 import api, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
d = scenes.boxroom_sequence(n)
dev = torch.from_numpy(d).cuda()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for streams in (1, 0):
    ext = api.PlaneExtractor(max_frames=n, n_streams=streams)
    ext.set_stream(stream.cuda_stream)
    ext.set_profile(streams == 1)
    for _ in range(3):
        ext.extract_device(dev.data_ptr(), n, 480, 640)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    reps = 5
    for _ in range(reps):
        ext.extract_device(dev.data_ptr(), n, 480, 640)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if streams == 1:
        kt = {}
        for name, t in ext.kernel_times():
            kt[name.split("<")[0]] = kt.get(name.split("<")[0], 0.0) + t
        top = sorted(kt.items(), key=lambda kv: -kv[1])
        print("one group: step %.3f ms | " % ms + " ".join(f"{k[2:]}={v:.3f}" for k, v in top[:12]))
    else:
        print("default groups: step %.3f ms -> %.0f frames/s" % (ms, n / ms * 1e3))
    ext.close()
print("env", {k: v for k, v in os.environ.items() if k.startswith("SPX_")})
