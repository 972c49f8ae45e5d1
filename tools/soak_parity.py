"""Parity soak beyond the test suite (development aid; needs a GPU and the oracle): every Frame field of N noisy box-room frames and of
M random clutter scenes against the oracle, bit for bit.   python tools/soak_parity.py [n_noisy] [n_clutter] > profiles/..."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sp_slam_b200 import api, scenes
from oracle import pyoracle

n_noisy = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_clutter = int(sys.argv[2]) if len(sys.argv) > 2 else 200
pyoracle.use_native()


def compare(res, frames, orc, tag):
    bad = []
    planes = points = 0
    for f in range(len(frames)):
        orc.run(frames[f])
        fp = res.frame(f)
        pr = orc.planes()
        ok = fp.mnRealPlaneNum == orc.n_real and fp.mnPlaneNum == orc.n_planes == len(pr)
        if ok:
            for i, b in enumerate(pr):
                ok = ok and fp.mvPlaneCoefficients[i].tobytes() == b["coef"].tobytes() and fp.mvPlanePoints[i].tobytes() == b["points"].tobytes() \
                    and fp.mvBoundaryPoints[i].tobytes() == b["boundary"].tobytes()
                points += len(b["points"])
        planes += len(pr)
        if not ok:
            bad.append(f)
    return {"set": tag, "frames": len(frames), "planes": planes, "cloud_points": points, "mismatching_frames": bad}


out = []
t0 = time.time()
d = scenes.boxroom_sequence(n_noisy)
d = np.stack([scenes.add_noise(d[k], k) for k in range(n_noisy)])
ext = api.PlaneExtractor(max_frames=max(n_noisy, n_clutter))
res = ext.extract_batch(d)
out.append(compare(res, d, pyoracle.Oracle(), "box-room orbit with sensor noise"))
if n_clutter:
    rng = np.random.default_rng(2024)
    frames = []
    for k in range(n_clutter):      # the box room plus a random clutter field, random pose; every third with sensor noise, some with dropouts
        rects = np.concatenate([scenes.boxroom_rects(), scenes.clutter_rects(n=int(rng.integers(5, 40)), seed=int(rng.integers(1 << 30)))])
        pose = scenes.poses(1000, seed=int(rng.integers(1 << 30)))[[int(rng.integers(1000))]]
        f = scenes.render(rects, pose, scenes.TUM1)[0]
        if k % 3 == 2:
            f = scenes.add_noise(f, k)
        if k % 4 == 1:
            f[rng.integers(0, 480, 40), rng.integers(0, 640, 40)] = 0.0
        frames.append(f)
    c = np.stack(frames)
    res = ext.extract_batch(c)
    out.append(compare(res, c, pyoracle.Oracle(), "box room + random clutter, random poses"))
print(json.dumps({"sets": out, "seconds": round(time.time() - t0, 1)}, indent=1))
