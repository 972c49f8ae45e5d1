"""End-to-end time of the host-input calls (1000 frames 640x480) under tuning knobs of the library (environment, read at spx_create).
python tools/e2e_knobs.py "K1=V1,K2=V2" "K3=V3" ...   (each argument = one configuration; "" = defaults)"""
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


for conf in sys.argv[1:] or [""]:
    env = dict(kv.split("=") for kv in conf.split(",") if kv)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        ext = api.PlaneExtractor(max_frames=F)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    out = []
    for mode in (2, 0, 3):
        ext.set_upload_mode(mode)
        out.append(f"mode {mode}: {timed(lambda: ext.extract_batch_compact_ptr(host.data_ptr(), F, rows, cols)):6.2f}")
    ext.set_upload_mode(0)
    out.append(f"u16: {timed(lambda: ext.extract_batch_u16_compact_ptr(host16.data_ptr(), F, rows, cols, factor)):6.2f}")
    dev = host.cuda()
    for _ in range(3):
        ext.extract_device(dev.data_ptr(), F, rows, cols)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ext.extract_device(dev.data_ptr(), F, rows, cols)
    torch.cuda.synchronize()
    out.append(f"resident: {(time.perf_counter() - t0) / 5 * 1e3:6.2f}")
    del dev
    print(f"[{conf or 'defaults'}]  " + "  ".join(out) + "  ms", flush=True)
    ext.close()
