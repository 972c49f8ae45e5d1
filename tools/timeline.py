"""Print the per-launch timeline of one resident step (development aid): python tools/timeline.py [frames] [streams]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sp_slam_b200 import api, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
st = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d = scenes.boxroom_sequence(n)
dev = torch.from_numpy(d).cuda()
ext = api.PlaneExtractor(max_frames=n, n_streams=st)
ext.set_profile(True)
for _ in range(4):
    ext.extract_device(dev.data_ptr(), n, 480, 640)
torch.cuda.synchronize()
tl = ext.kernel_timeline()
seen = {}
for name, a, b in tl:
    g = seen.get(name, 0); seen[name] = g + 1
    print(f"{name.split('<')[0]:18s} g{g:<2d} {a:7.3f} -> {b:7.3f}  ({b - a:6.3f})")
print("end", max(b for _, _, b in tl))
