"""Wall time of the host-input call (development aid): python tools/e2e_wall.py [frames] [streams] [mode] [u16]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sp_slam_b200 import api, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
st = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
u16 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
compact = int(sys.argv[5]) if len(sys.argv) > 5 else 0
d = scenes.boxroom_sequence(n)
if u16:
    d = np.round(np.clip(d, 0, 13.0).astype(np.float64) * 5000.0).astype(np.uint16)
host = torch.from_numpy(d).pin_memory()
ext = api.PlaneExtractor(max_frames=n, n_streams=st)
ext.set_upload_mode(mode)
f = float(np.float32(1.0) / np.float32(5000.0))
if compact:
    call = (lambda: ext.extract_batch_u16_compact_ptr(host.data_ptr(), n, 480, 640, f)) if u16 else (lambda: ext.extract_batch_compact_ptr(host.data_ptr(), n, 480, 640))
else:
    call = (lambda: ext.extract_batch_u16_ptr(host.data_ptr(), n, 480, 640, f)) if u16 else (lambda: ext.extract_batch_ptr(host.data_ptr(), n, 480, 640))
for _ in range(3):
    call()
ts = []
for _ in range(10):
    t0 = time.perf_counter(); call(); ts.append((time.perf_counter() - t0) * 1e3)
print(f"n={n} streams={st} mode={mode} u16={u16} compact={compact} xfer={ext.transfer_bytes()} env={ {k: v for k, v in os.environ.items() if k.startswith('SPX_')} }: median {np.median(ts):.2f} min {min(ts):.2f} ms -> {n / np.median(ts) * 1e3:.0f} frames/s")
np.set_printoptions(precision=2, suppress=True, linewidth=200)
print("group timeline (start, uploaded, postfilter, kernels end, on host):")
try:
    print(ext.group_timeline())
except BrokenPipeError:
    pass
