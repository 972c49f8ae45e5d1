"""Gathered upload (spx_set_upload_mode 3) against the sampled-rows upload (mode 2) on the box it runs on:
(1) the host gather alone, per thread count; (2) spx_extract_batch_compact end to end, 1000 frames 640x480, per mode /
thread count / group count.  python tools/gather_probe.py [frames]"""
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sp_slam_b200 import api, scenes

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
d = scenes.boxroom_sequence(F)
host = torch.from_numpy(d).pin_memory()
rows, cols = d.shape[1:]
L = api.lib()
print("cpus", os.cpu_count(), flush=True)
out = torch.empty((F, 160, 216), dtype=torch.float32).pin_memory()
for t in (1, 2, 4, 8, 12, 16, 24, 32):
    if t > (os.cpu_count() or 1):
        break
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        L.spx_host_gather_samples(host.data_ptr(), F, rows, cols, cols * 4, cols * rows * 4, 3, 8, t, out.data_ptr(), 216)
        best = min(best, time.perf_counter() - t0)
    print(f"gather alone: {t:2d} threads {best * 1e3 * 1000 / F:7.2f} ms per 1000 frames ({F * 160 * cols * 4 / best / 1e9:6.1f} GB/s of sampled rows)", flush=True)


def timed(ext, steps=6):
    for _ in range(2):
        ext.extract_batch_compact_ptr(host.data_ptr(), F, rows, cols)
    t0 = time.perf_counter()
    for _ in range(steps):
        ext.extract_batch_compact_ptr(host.data_ptr(), F, rows, cols)
    return (time.perf_counter() - t0) / steps * 1e3


ncpu = os.cpu_count() or 1
for ns in (8, 12):
    ext = api.PlaneExtractor(max_frames=F, n_streams=ns)
    ext.set_upload_mode(2)
    print(f"groups {ns:2d}  sampled rows only: {timed(ext):6.2f} ms", flush=True)
    ext.set_upload_mode(3)
    for t in (8, 12):
        if t <= ncpu:
            ext.set_gather_threads(t)
            print(f"groups {ns:2d}  all gathered, {t:2d} threads: {timed(ext):6.2f} ms  h2d {ext.transfer_bytes()[0] / 1e6:.1f} MB", flush=True)
    ext.set_upload_mode(0)
    for t in (6, 8, 10, 12):
        if t > ncpu:
            break
        ext.set_gather_threads(t)
        line = []
        for share in (0.3, 0.4, 0.5, 0.6, 0.7):
            ext.set_gather_share(share)
            line.append(f"{share:.1f}: {timed(ext):6.2f} ms ({ext.transfer_bytes()[0] / 1e6:.0f} MB)")
        print(f"groups {ns:2d}  both routes, {t:2d} threads, share " + "  ".join(line), flush=True)
    ext.set_gather_threads(8)
    ext.set_gather_share(0.5)
    timed(ext, 2)
    print("   timeline at 8 threads, share 0.5 (start, uploaded, real planes, end, on host | host enqueue, host totals):")
    for row in ext.group_timeline():
        print("   ", " ".join(f"{v:6.2f}" for v in row))
    ext.close()
