"""Shared helpers of the SURVEY 8(f) tests: point construction and independent numpy restatements (fp32 op by op)."""
import numpy as np

POINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])


def make_points(xyz, argb=None):
    xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
    p = np.zeros(len(xyz), POINT)
    p["x"], p["y"], p["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    if argb is None:
        p["rgba"] = (255 << 24) | 250
    else:
        c = np.asarray(argb, np.uint32).reshape(-1, 4)
        p["rgba"] = (c[:, 0] << 24) | (c[:, 1] << 16) | (c[:, 2] << 8) | c[:, 3]
    return p


def np_voxel_groups(pts, leaf):
    """voxel index of every finite point as PCL 1.8.0 VoxelGrid computes it; returns (sorted unique keys, members per key)"""
    f = np.float32
    leaf3 = np.broadcast_to(np.asarray(leaf, f), (3,))
    inv = f(1.0) / leaf3
    xyz = np.stack([pts["x"], pts["y"], pts["z"]], 1)
    fin = np.isfinite(xyz).all(1)
    src = np.nonzero(fin)[0]
    xyz = xyz[fin]
    mn, mx = xyz.min(0), xyz.max(0)
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div = max_b - min_b + 1
    ijk = (np.floor(xyz * inv) - min_b.astype(f)).astype(np.int32)
    key = ijk[:, 0].astype(np.int64) + ijk[:, 1].astype(np.int64) * div[0] + ijk[:, 2].astype(np.int64) * div[0] * div[1]
    order = np.argsort(key, kind="stable")
    keys, first = np.unique(key[order], return_index=True)
    bounds = list(first) + [len(order)]
    return keys, [src[order[bounds[k]:bounds[k + 1]]] for k in range(len(keys))]


def np_associate(plane_w, map_w, bnds, n_seen, dis_th=0.2, ang_th=0.8, ver_th=0.08716, par_th=0.9962):
    f = np.float32
    n = len(plane_w)
    a, v, p = (np.full(n, -1, np.int32) for _ in range(3))
    d = np.zeros(n, f)

    def dist(pl, b):
        if len(b) == 0:
            return 100.0
        e = ((pl[0] * b["x"] + pl[1] * b["y"]) + pl[2] * b["z"]) + pl[3]      # fp32, left to right
        e = np.abs(e[~np.isnan(e)])
        return min(100.0, float(e.min())) if len(e) else 100.0

    for i in range(n):
        pl = plane_w[i].astype(f)
        ld, lv, lp = f(dis_th), f(ver_th), f(par_th)
        for j in range(len(map_w)):
            w = map_w[j].astype(f)
            ang = (pl[0] * w[0] + pl[1] * w[1]) + pl[2] * w[2]
            seen = j < n_seen
            if not seen and not (ld == f(dis_th) or a[i] >= n_seen):
                break
            if ang > f(ang_th) or ang < -f(ang_th):
                dd = dist(pl, bnds[j])
                if dd < ld:
                    ld = f(dd); a[i] = j
                    continue
            if not seen:
                continue
            if -lv < ang < lv:
                lv = abs(ang); v[i] = j
                continue
            if ang > lp or ang < -lp:
                lp = abs(ang); p[i] = j
        d[i] = ld
    return a, v, p, d
