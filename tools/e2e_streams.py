"""End-to-end time of the host-input calls (1000 frames 640x480) against the number of frame groups.  python tools/e2e_streams.py 2 3 4 8"""
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from sp_slam_b200 import api, scenes

F = 1000
d = scenes.boxroom_sequence(F)
host = torch.from_numpy(d).pin_memory()
host16 = torch.from_numpy(np.round(np.clip(d, 0, 13.0).astype(np.float64) * 5000.0).astype(np.uint16)).pin_memory()
factor = float(np.float32(1.0) / np.float32(5000.0))
rows, cols = d.shape[1:]


def timed(call, steps=6):
    for _ in range(2):
        call()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    return (time.perf_counter() - t0) / steps * 1e3


for ns in [int(a) for a in sys.argv[1:]] or [2, 3, 4, 8]:
    ext = api.PlaneExtractor(max_frames=F, n_streams=ns)
    out = []
    for mode in (2, 0):
        ext.set_upload_mode(mode)
        out.append(f"mode {mode}: {timed(lambda: ext.extract_batch_compact_ptr(host.data_ptr(), F, rows, cols)):6.2f}")
    out.append(f"u16: {timed(lambda: ext.extract_batch_u16_compact_ptr(host16.data_ptr(), F, rows, cols, factor)):6.2f}")
    print(f"groups {ns:2d}  " + "  ".join(out) + "  ms", flush=True)
    if ns <= 4:
        for row in ext.group_timeline():
            print("      ", " ".join(f"{v:6.2f}" for v in row))
    ext.close()
