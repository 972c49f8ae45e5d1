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


def next_rows_fixture():
    """inputs and oracle outputs of the SURVEY 8(f) rows: VoxelGrid (N4), plane association + boundary transform (N1)"""
    rng = np.random.default_rng(2024)
    POINT = pyoracle.POINT_DTYPE
    def cloud(n, spread):
        p = np.zeros(n, POINT)
        xyz = (rng.normal(size=(n, 3)) * spread + [0.2, -0.1, 2.0]).astype(np.float32)
        p["x"], p["y"], p["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
        c = rng.integers(0, 256, size=(n, 4)).astype(np.uint32)
        p["rgba"] = (c[:, 0] << 24) | (c[:, 1] << 16) | (c[:, 2] << 8) | c[:, 3]
        return p
    vox_in = cloud(3000, [0.25, 0.15, 0.03])
    vox = {}
    for leaf in (0.01, 0.05):
        out, idx = pyoracle.voxel_grid(vox_in, leaf)
        vox[f"vox_out_{leaf}"] = out
        vox[f"vox_idx_{leaf}"] = idx
    n_map, n_pl = 9, 6
    nrm = np.eye(3)[rng.integers(0, 3, n_map)] * rng.choice([-1, 1], (n_map, 1)) + 0.03 * rng.normal(size=(n_map, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    map_w = np.concatenate([nrm, rng.uniform(-3, 3, (n_map, 1))], 1).astype(np.float32)
    sizes = rng.integers(0, 400, n_map)
    bnd = cloud(int(sizes.sum()), [3.0, 3.0, 3.0])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    pn = np.eye(3)[rng.integers(0, 3, n_pl)] + 0.03 * rng.normal(size=(n_pl, 3))
    pn /= np.linalg.norm(pn, axis=1, keepdims=True)
    plane_w = np.concatenate([pn, rng.uniform(-3, 3, (n_pl, 1))], 1).astype(np.float32)
    bnds = [bnd[off[j]:off[j + 1]] for j in range(n_map)]
    a, v, p, d = pyoracle.associate_planes(plane_w, map_w, bnds, n_seen=6)
    m = np.eye(4)
    m[:3, :] = rng.normal(size=(3, 4))
    return dict(vox_in=vox_in, **vox, map_w=map_w, map_bnd=bnd, map_off=off, plane_w=plane_w, n_seen=6, assoc=a, vertical=v,
                parallel=p, assoc_dist=d, xform=m, xform_out=pyoracle.transform_cloud(bnds[int(np.argmax(sizes))], m),
                xform_src=int(np.argmax(sizes)))


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
    np.savez_compressed(os.path.join(out, "next_rows.npz"), **next_rows_fixture())


if __name__ == "__main__":
    main()
