"""Host-to-host throughput with several batches in flight: T host threads, each with its own context and its own pinned batch,
all calling spx_extract_batch_compact (development aid): python tools/e2e_two.py frames threads u16 steps"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sp_slam_b200 import api, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2
u16 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 12
d = scenes.boxroom_sequence(n)
if u16:
    d = np.round(np.clip(d, 0, 13.0).astype(np.float64) * 5000.0).astype(np.uint16)
f = float(np.float32(1.0) / np.float32(5000.0))
hosts = [torch.from_numpy(d.copy()).pin_memory() for _ in range(T)]
exts = [api.PlaneExtractor(max_frames=n) for _ in range(T)]
def call(t):
    if u16: return exts[t].extract_batch_u16_compact_ptr(hosts[t].data_ptr(), n, 480, 640, f)
    return exts[t].extract_batch_compact_ptr(hosts[t].data_ptr(), n, 480, 640)
for t in range(T):
    for _ in range(3): call(t)
start = threading.Barrier(T + 1)
def work(t, k):
    start.wait()
    for _ in range(k): call(t)
best = 1e9
for rep in range(3):
    per = steps // T
    th = [threading.Thread(target=work, args=(t, per)) for t in range(T)]
    for x in th: x.start()
    start.wait(); t0 = time.perf_counter()
    for x in th: x.join()
    dt = (time.perf_counter() - t0) * 1e3 / (per * T)
    best = min(best, dt)
print(f"threads {T} u16 {u16}: {best:.2f} ms per batch -> {n / best * 1e3:.0f} frames/s")
