"""Two contexts, two batches in flight (development aid): python tools/two_ctx.py frames contexts"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sp_slam_b200 import api, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = scenes.boxroom_sequence(n)
dev = torch.from_numpy(d).cuda()
streams = [torch.cuda.Stream() for _ in range(nctx)]
exts = []
for s in streams:
    e = api.PlaneExtractor(max_frames=n)
    e.set_stream(s.cuda_stream)
    exts.append(e)
for k in range(3 * nctx):
    exts[k % nctx].extract_device(dev.data_ptr(), n, 480, 640)
torch.cuda.synchronize()
best = 1e9
reps = 12
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams: s.wait_event(e0)
    for k in range(reps):
        exts[k % nctx].extract_device(dev.data_ptr(), n, 480, 640)
    for s in streams:
        ev = torch.cuda.Event(); ev.record(s); main.wait_event(ev)
    e1.record(main)
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / reps)
print("contexts %d: %.3f ms per batch -> %.0f frames/s" % (nctx, best, n / best * 1e3))
