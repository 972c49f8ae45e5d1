"""Small fixed workload for ncu captures of the host-input path: N box-room frames from a page-locked buffer, 3 passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sp_slam_b200 import api, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250
d = scenes.boxroom_sequence(n, start=150)
host = torch.from_numpy(d).pin_memory()
ext = api.PlaneExtractor(max_frames=n, n_streams=1)
for _ in range(3):
    r = ext.extract_batch_ptr(host.data_ptr(), n, 480, 640)
print("frames", n, "planes", int(r.frames["n_planes"].sum()), "transfer", ext.transfer_bytes())
