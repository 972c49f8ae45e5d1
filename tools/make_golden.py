"""Generate tests/golden/*.npz from the CPU oracle (run in the build container; the fixtures are committed).

The reference ships no golden vectors for this path and its PCL dependency cannot be built here, so these fixtures
freeze the ORACLE's outputs (parity unpinned, see oracle/spx_oracle.h).  They guard against drift of the oracle, the
scene renderer and the noise models, and give the GPU tests fixed targets that do not depend on rebuilding anything.

    python tools/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from sp_slam_b200 import scenes  # noqa: E402

CASES = [("boxroom_f200_clean", 200, False), ("boxroom_f880_clean", 880, False), ("boxroom_f640_kinect", 640, True)]


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def summarize(o):
    pls = o.planes()
    recs = o.line_recs()
    return dict(
        n_real=o.n_real, n_planes=o.n_planes,
        coef=np.stack([p["coef"] for p in pls]) if pls else np.zeros((0, 4), np.float32),
        n_points=np.array([len(p["points"]) for p in pls], np.int32),
        n_boundary=np.array([len(p["boundary"]) for p in pls], np.int32),
        src=np.array([p["src"] for p in pls], np.int32),
        points_sha=np.array([sha(p["points"]) for p in pls]),
        boundary_sha=np.array([sha(p["boundary"]) for p in pls]),
        labels_raw_sha=sha(o.labels_raw()[0]), n_label_lists=o.labels_raw()[1],
        labels_refined_sha=sha(o.labels_refined()),
        normals_sha=sha(o.normals()), distance_sha=sha(np.minimum(o.distance_map(), np.float32(10))),
        model_labels=np.array([m["label"] for m in o.models()], np.int64),
        line_recs=np.array([[r[k] for k in ("plane", "round", "n_points", "iterations", "n_inliers", "in_range",
                                            "is_border", "emitted")] for r in recs], np.int32).reshape(-1, 8),
        line_coef=np.array([r["coef"] for r in recs], np.float32).reshape(-1, 6),
    )


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    P = scenes.poses(1000)
    for name, f, noisy in CASES:
        d = scenes.render(scenes.boxroom_rects(), P[[f]], scenes.TUM1)[0]
        extra = {}
        if noisy:
            d = scenes.add_noise(d, f)
            extra["depth_u16"] = np.round(d.astype(np.float64) * 5000.0).astype(np.uint16)   # exact: d = u16 / 5000
        o = pyoracle.Oracle().run(d)
        np.savez_compressed(os.path.join(out, name + ".npz"), frame=f, noisy=noisy, depth_sha=sha(d), **extra,
                            **summarize(o))
        print(name, "planes", o.n_real, o.n_planes, "depth", sha(d)[:12])


if __name__ == "__main__":
    main()
