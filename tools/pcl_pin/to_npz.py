#!/usr/bin/env python
"""Packs the .npy dumps of tools/pcl_pin/pcl_pin (one directory per frame) together with the frame's depth image into
tests/golden/pcl_<frame>.npz, the files tests/test_pcl_pin.py looks for.
    python tools/pcl_pin/to_npz.py IN_DIR(make_inputs) OUT_DIR(pcl_pin dumps) tests/golden"""
import json
import os
import sys

import numpy as np

KEYS = ("dims", "cloud", "cloud_rgba", "normals", "seg_labels", "seg_n_label_lists", "seg_coef", "seg_inliers_val", "seg_inliers_off",
        "seg_centroids", "seg_cov", "ref_labels", "ref_coef", "ref_inliers_val", "ref_inliers_off", "ref_boundary_idx_val",
        "ref_boundary_idx_off", "contour_pts", "contour_off", "line_rec", "line_coef", "line_inliers_val", "line_inliers_off")


def pack(depth, meta, dump_dir, dst, source):
    arrays = {k: np.load(os.path.join(dump_dir, k + ".npy")) for k in KEYS}
    np.savez_compressed(dst, depth=depth, intrinsics=np.array([meta["fx"], meta["fy"], meta["cx"], meta["cy"]], np.float32),
                        source=np.array(source), **arrays)


def main():
    in_dir, out_dir, golden = sys.argv[1:4]
    manifest = json.load(open(os.path.join(in_dir, "manifest.json")))
    for name, meta in manifest.items():
        d = os.path.join(out_dir, name)
        if not os.path.isdir(d):
            print("skipping", name, "(no dump directory)")
            continue
        depth = np.fromfile(os.path.join(in_dir, name + ".bin"), np.float32).reshape(meta["rows"], meta["cols"])
        pack(depth, meta, d, os.path.join(golden, f"pcl_{name}.npz"), "PCL (tools/pcl_pin/pcl_pin)")
        print("wrote", os.path.join(golden, f"pcl_{name}.npz"))


if __name__ == "__main__":
    main()
