"""Small fixed workload for ncu captures: N box-room frames through the device-resident entry point, 3 passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sp_slam_b200 import api, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
d = scenes.boxroom_sequence(n, start=150)
dev = torch.from_numpy(d).cuda()
st = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # frame groups (internal streams); 1 = every kernel once per pass
ext = api.PlaneExtractor(max_frames=n, n_streams=st)
for _ in range(3):
    ext.extract_device(dev.data_ptr(), n, 480, 640)
r = ext.fetch(clouds=False)
print("frames", n, "planes", int(r.frames["n_planes"].sum()))
