#!/usr/bin/env python
"""Exactness of the fp64 integral images over the bench workloads.

The CUDA path builds tile-local integral images; they give bit-identical window sums to PCL's whole-image ones only
if every fp64 partial sum of the fp32 central differences is exact.  The oracle checks that with error-free
transformations (Sat3::add_chk).  This sweeps all 1000 box-room frames, clean and with sensor noise (float and the
16-bit PNG encoding), and the 1280x720 clutter sequence, and writes profiles/r2_sat_exact_sweep.json.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from sp_slam_b200 import scenes  # noqa: E402


def main():
    threads = os.cpu_count() or 1
    out = {"threads": threads, "sets": []}
    t0 = time.time()
    clean = scenes.boxroom_sequence(1000)
    sets = [("boxroom 640x480 clean, frames 0..999", clean, pyoracle.default_config())]
    noisy = np.stack([scenes.add_noise(clean[k], k, "kinect") for k in range(1000)])
    sets.append(("boxroom 640x480 kinect noise + 2% dropouts (float)", noisy, pyoracle.default_config()))
    q16 = (np.round(np.clip(noisy, 0, 13.0).astype(np.float64) * 5000.0).astype(np.uint16).astype(np.float32)
           * np.float32(np.float32(1.0) / np.float32(5000.0)))
    sets.append(("the same through the 16-bit PNG encoding (DepthMapFactor 5000)", q16, pyoracle.default_config()))
    it = scenes.REALSENSE
    big = scenes.realsense_sequence(200)
    big = np.stack([scenes.add_noise(big[k], k, "realsense") for k in range(len(big))])
    sets.append(("realsense 1280x720 clutter + noise, frames 0..199", big,
                 pyoracle.default_config(fx=it.fx, fy=it.fy, cx=it.cx, cy=it.cy, max_x=float(it.width), max_y=float(it.height))))
    for name, d, cfg in sets:
        ex = pyoracle.sat_exact_batch(d, threads, cfg)
        out["sets"].append({"set": name, "frames": int(len(d)), "exact_frames": int(ex.sum()),
                            "inexact_frames": [int(i) for i in np.nonzero(~ex)[0][:50]]})
        print(out["sets"][-1], flush=True)
    out["seconds"] = time.time() - t0
    with open(os.path.join(ROOT, "profiles", "r2_sat_exact_sweep.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
