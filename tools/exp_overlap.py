"""Does a concurrent H2D / D2H copy slow the resident kernels down? (development aid)"""
import os, sys, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sp_slam_b200 import api, scenes
n = 250
d = scenes.boxroom_sequence(n)
host = torch.from_numpy(d).pin_memory()
dev = host.cuda()
big_h = torch.empty(1 << 28, dtype=torch.float32).pin_memory()   # 1 GiB
big_d = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
ext = api.PlaneExtractor(max_frames=n, n_streams=1)
s_k, s_c, s_c2 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
ext.set_stream(s_k.cuda_stream)
def run(copy_h2d, copy_d2h, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        if copy_h2d:
            with torch.cuda.stream(s_c):
                big_d.copy_(big_h, non_blocking=True)
        if copy_d2h:
            with torch.cuda.stream(s_c2):
                big_h.copy_(big_d, non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s_k):
            e0.record(s_k)
            ext.extract_device(dev.data_ptr(), n, 480, 640)
            e1.record(s_k)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return np.median(ts)
for _ in range(3):
    ext.extract_device(dev.data_ptr(), n, 480, 640)
print("alone", run(False, False))
print("with H2D", run(True, False))
print("with D2H", run(False, True))
print("with both", run(True, True))
q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,pstate,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout
print("idle clocks", q.strip())
# clocks during an e2e loop
ext2 = api.PlaneExtractor(max_frames=n, n_streams=4)
stop = False
samples = []
def sampler():
    while not stop:
        samples.append(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,utilization.gpu", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
th = threading.Thread(target=sampler); th.start()
t0 = time.time()
while time.time() - t0 < 3:
    ext2.extract_batch_ptr(host.data_ptr(), n, 480, 640)
stop = True; th.join()
print("clocks during e2e loop", samples[::max(1, len(samples) // 8)])
