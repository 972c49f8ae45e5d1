#!/usr/bin/env python
"""Writes the depth frames tools/pcl_pin/pcl_pin is to be run on, as raw float32 (rows x cols, metres):
the three fixture frames of tests/golden (box room 200 and 880 clean, 640 with Kinect noise through the 16-bit PNG
encoding) and one 1280x720 RealSense-shaped clutter frame.      python tools/pcl_pin/make_inputs.py OUT_DIR"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sp_slam_b200 import scenes  # noqa: E402


def frames():
    P = scenes.poses(1000)
    out = {}
    for name, f, noisy in (("boxroom_f200_clean", 200, False), ("boxroom_f880_clean", 880, False), ("boxroom_f640_kinect", 640, True)):
        d = scenes.render(scenes.boxroom_rects(), P[[f]], scenes.TUM1)[0]
        if noisy:
            d = scenes.add_noise(d, f, "kinect")
        out[name] = (d, scenes.TUM1)
    big = scenes.add_noise(scenes.realsense_sequence(1, start=40)[0], 40, "realsense")
    out["realsense_f40_noisy"] = (big, scenes.REALSENSE)
    return out


def main():
    out_dir = sys.argv[1]
    os.makedirs(out_dir, exist_ok=True)
    manifest = {}
    for name, (d, it) in frames().items():
        np.ascontiguousarray(d, np.float32).tofile(os.path.join(out_dir, name + ".bin"))
        manifest[name] = {"rows": int(d.shape[0]), "cols": int(d.shape[1]), "fx": it.fx, "fy": it.fy, "cx": it.cx, "cy": it.cy}
        print(f"pcl_pin {out_dir}/{name}.bin {d.shape[0]} {d.shape[1]} OUT/{name} {it.fx!r} {it.fy!r} {it.cx!r} {it.cy!r}")
    json.dump(manifest, open(os.path.join(out_dir, "manifest.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
