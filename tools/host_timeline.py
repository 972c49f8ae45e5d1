"""Where the device time of a host-input call goes: per kernel, the sum over the frame groups of (end - start) from the
library's launch timeline (events around every launch: spx_set_profile), beside the same kernel's time in a resident
1000-frame single-group step; and how many kernels are in flight over time.  python tools/host_timeline.py [u16|f32]"""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sp_slam_b200 import api, scenes

kind = sys.argv[1] if len(sys.argv) > 1 else "u16"
F = 1000
d = scenes.boxroom_sequence(F)
rows, cols = d.shape[1:]
host = torch.from_numpy(d).pin_memory()
host16 = torch.from_numpy(np.round(np.clip(d, 0, 13.0).astype(np.float64) * 5000.0).astype(np.uint16)).pin_memory()
factor = float(np.float32(1.0) / np.float32(5000.0))

one = api.PlaneExtractor(max_frames=F, n_streams=1)
one.set_profile(True)
dev = host.cuda()
for _ in range(3):
    one.extract_device(dev.data_ptr(), F, rows, cols)
torch.cuda.synchronize()
base = {}
for name, t in one.kernel_times():
    base[name.split("<")[0]] = base.get(name.split("<")[0], 0.0) + t
one.close()
del dev

ext = api.PlaneExtractor(max_frames=F)
ext.set_profile(True)
call = (lambda: ext.extract_batch_u16_compact_ptr(host16.data_ptr(), F, rows, cols, factor)) if kind == "u16" else \
       (lambda: ext.extract_batch_compact_ptr(host.data_ptr(), F, rows, cols))
for _ in range(3):
    call()
tl = ext.kernel_timeline()
gt = ext.group_timeline()
tot = {}
for name, a, b in tl:
    k = name.split("<")[0]
    tot[k] = tot.get(k, 0.0) + (b - a)
end = max(b for _, a, b in tl)
print(f"{kind}: {len(tl)} launches, last kernel ends at {end:.2f} ms, {len(gt)} groups")
print("kernel                     sum over groups   resident single group   ratio")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    b = base.get(k)
    print(f"{k:26s} {v:10.3f} {b if b is not None else float('nan'):18.3f} {v / b if b else float('nan'):12.2f}")
print(f"sum {sum(tot.values()):.2f} ms of kernel spans; resident single group {sum(base.values()):.2f} ms")
# kernels in flight per 0.25 ms
step = 0.25
print("t(ms)  kernels in flight (average), names")
t = 0.0
while t < end:
    live = [(n_.split('<')[0][2:], max(a, t), min(b, t + step)) for n_, a, b in tl if b > t and a < t + step]
    avg = sum(hi - lo for _, lo, hi in live) / step
    names = {}
    for n_, lo, hi in live:
        names[n_] = names.get(n_, 0.0) + (hi - lo) / step
    print(f"{t:5.2f}  {avg:5.2f}  " + " ".join(f"{k}:{v:.1f}" for k, v in sorted(names.items(), key=lambda kv: -kv[1])[:7]))
    t += step
print("groups (start, uploaded, real planes, end, on host | host enqueue, host totals):")
for row in gt:
    print("   ", " ".join(f"{v:6.2f}" for v in row))
