import os

import numpy as np

from sp_slam_b200 import scenes

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["boxroom_f200_clean", "boxroom_f880_clean", "boxroom_f640_kinect"]


def load(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    if "depth_u16" in g.files:
        depth = (g["depth_u16"].astype(np.float32) * np.float32(1.0 / 5000.0)).astype(np.float32)
    else:
        depth = scenes.render(scenes.boxroom_rects(), scenes.poses(1000)[[int(g["frame"])]], scenes.TUM1)[0]
    return g, depth
