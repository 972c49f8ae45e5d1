"""ctypes binding of the CPU oracle (oracle/libspx_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product package.  PARITY UNPINNED (see spx_oracle.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libspx_oracle.so")


class OrcConfig(C.Structure):
    _fields_ = [
        ("cloud_dis", C.c_int32), ("min_size", C.c_int32), ("angle_thr_deg", C.c_float), ("dist_thr", C.c_float),
        ("line_ratio", C.c_double), ("line_dist_thr", C.c_float),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
        ("max_depth_change_factor", C.c_float), ("normal_smoothing_size", C.c_float),
        ("ransac_max_iter", C.c_int32), ("enable_supposed", C.c_int32),
        ("alt", C.c_uint32),      # alternative readings of PCL 1.8.0 (ORC_ALT_*), 0 = the default restatement
        ("normal_method", C.c_int32),   # 0 AVERAGE_3D_GRADIENT (the reference), 1 COVARIANCE_MATRIX
    ]


ALT_VP_RESET, ALT_CHAMFER_NO_WRAP, ALT_REFINE_NO_WRAP, ALT_SAMPLE_GOOD_OR, ALT_RNG_MASK, ALT_SO_DOUBLE, ALT_COV_TRACE = 1, 2, 4, 8, 16, 32, 64


class OrcLineRec(C.Structure):
    _fields_ = [
        ("plane", C.c_int32), ("round", C.c_int32), ("n_points", C.c_int32), ("iterations", C.c_int32),
        ("n_inliers", C.c_int32), ("in_range", C.c_int32), ("is_border", C.c_int32), ("emitted", C.c_int32),
        ("coef", C.c_float * 6),
    ]


POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("spx_oracle.cpp", "spx_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libspx_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_lib_path = _LIB_PATH
BUILD_FLAGS = "-O3 -ffp-contract=off (portable build)"


def use_native() -> str:
    """Rebuild the oracle for THIS host (-O3 -march=native -ffp-contract=off, the reference's CMakeLists.txt:10-11 flags plus
    the no-contraction rule) into oracle/_native/ and use it from now on; bench.py's CPU legs call this on the box they run
    on.  Falls back to the portable build (and says so) when the compiler is missing.  Returns the flags in use."""
    global _lib, _lib_path, BUILD_FLAGS
    if _lib is not None:
        return BUILD_FLAGS
    out_dir = os.path.join(_HERE, "_native")
    out = os.path.join(out_dir, "libspx_oracle_native.so")
    try:
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared",
                               "-pthread", "-o", out, os.path.join(_HERE, "spx_oracle.cpp")], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL, timeout=300)
        _lib_path = out
        BUILD_FLAGS = "-O3 -march=native -ffp-contract=off (built on this host)"
    except Exception as e:   # noqa: BLE001
        BUILD_FLAGS = f"-O3 -ffp-contract=off (portable build; the native rebuild failed: {type(e).__name__})"
    return BUILD_FLAGS


def lib():
    global _lib
    if _lib is None:
        if _lib_path == _LIB_PATH and not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_lib_path)
        vp, i32, f32p = C.c_void_p, C.c_int, C.c_void_p
        L.orc_default_config.argtypes = [C.POINTER(OrcConfig)]
        L.orc_create.argtypes = [C.POINTER(OrcConfig)]
        L.orc_create.restype = vp
        L.orc_destroy.argtypes = [vp]
        L.orc_run.argtypes = [vp, f32p, i32, i32, f32p]
        L.orc_run.restype = i32
        L.orc_dims.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
        for name in ("orc_get_cloud", "orc_get_normals"):
            getattr(L, name).argtypes = [vp, vp, vp, vp]
        for name in ("orc_get_distance_map", "orc_get_plane_d", "orc_get_labels_refined", "orc_get_curvature"):
            getattr(L, name).argtypes = [vp, vp]
        L.orc_get_labels_raw.argtypes = [vp, vp]
        L.orc_get_labels_raw.restype = i32
        for name in ("orc_num_models", "orc_sat_exact", "orc_num_real_planes", "orc_num_planes", "orc_num_line_recs"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = i32
        L.orc_get_model.argtypes = [vp, i32, vp, vp, vp, C.POINTER(C.c_float), C.POINTER(C.c_uint32),
                                    C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.orc_get_model_inliers.argtypes = [vp, i32, vp]
        L.orc_get_model_contour.argtypes = [vp, i32, vp]
        L.orc_get_plane.argtypes = [vp, i32, vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.orc_get_plane_points.argtypes = [vp, i32, vp]
        L.orc_get_plane_boundary.argtypes = [vp, i32, vp]
        L.orc_get_line_recs.argtypes = [vp, vp]
        L.orc_get_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_run_batch.argtypes = [C.POINTER(OrcConfig), vp, i32, i32, i32, i32, vp, vp, C.POINTER(C.c_double),
                                    C.POINTER(C.c_double)]
        L.orc_run_batch.restype = i32
        L.orc_chamfer.argtypes = [vp, i32, i32, vp]
        L.orc_eigen33_smallest.argtypes = [vp, C.POINTER(C.c_float), vp]
        L.orc_eigen33_largest.argtypes = [vp, vp, vp]
        L.orc_sac_line.argtypes = [vp, i32, C.c_double, i32, vp, vp, C.POINTER(i32)]
        L.orc_sac_line.restype = i32
        L.orc_sac_line_alt.argtypes = [vp, i32, C.c_double, i32, vp, vp, C.POINTER(i32), C.c_uint32]
        L.orc_sac_line_alt.restype = i32
        L.orc_chamfer_alt.argtypes = [vp, i32, i32, vp, C.c_uint32]
        L.orc_is_border_point.argtypes = [C.POINTER(OrcConfig), vp, i32, i32, C.c_float, C.c_float, C.c_float]
        L.orc_is_border_point.restype = i32
        L.orc_sat_exact_batch.argtypes = [C.POINTER(OrcConfig), vp, i32, i32, i32, i32, vp]
        L.orc_ransac_draws.argtypes = [i32, i32, vp]
        L.orc_voxel_grid.argtypes = [vp, i32, vp, vp, vp]
        L.orc_voxel_grid.restype = i32
        L.orc_associate_planes.argtypes = [vp, i32, vp, vp, vp, i32, i32, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp]
        L.orc_associate_planes.restype = None
        L.orc_transform_cloud.argtypes = [vp, i32, vp, vp]
        L.orc_transform_cloud.restype = None
        _lib = L
    return _lib


def default_config(**overrides) -> OrcConfig:
    cfg = OrcConfig()
    lib().orc_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


class Oracle:
    """One oracle context; ``run`` executes the whole path on a frame and keeps every intermediate."""

    def __init__(self, cfg: OrcConfig | None = None, **overrides):
        self.cfg = cfg if cfg is not None else default_config(**overrides)
        self._h = lib().orc_create(C.byref(self.cfg))
        self.width = self.height = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def run(self, depth: np.ndarray, normals: np.ndarray | None = None) -> "Oracle":
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        rows, cols = depth.shape
        nptr = None
        if normals is not None:
            normals = np.ascontiguousarray(normals, dtype=np.float32)
            nptr = normals.ctypes.data
        self._depth = depth
        lib().orc_run(self._h, depth.ctypes.data, rows, cols, nptr)
        w, h = C.c_int(), C.c_int()
        lib().orc_dims(self._h, C.byref(w), C.byref(h))
        self.width, self.height = w.value, h.value
        return self

    @property
    def n(self) -> int:
        return self.width * self.height

    def _three(self, fn):
        a = [np.empty(self.n, np.float32) for _ in range(3)]
        fn(self._h, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data)
        return np.stack(a)

    def cloud(self):
        return self._three(lib().orc_get_cloud)

    def normals(self):
        return self._three(lib().orc_get_normals)

    def distance_map(self):
        d = np.empty(self.n, np.float32)
        lib().orc_get_distance_map(self._h, d.ctypes.data)
        return d.reshape(self.height, self.width)

    def curvature(self):
        d = np.empty(self.n, np.float32)
        lib().orc_get_curvature(self._h, d.ctypes.data)
        return d

    def plane_d(self):
        d = np.empty(self.n, np.float32)
        lib().orc_get_plane_d(self._h, d.ctypes.data)
        return d

    def labels_raw(self):
        l = np.empty(self.n, np.uint32)
        n_lists = lib().orc_get_labels_raw(self._h, l.ctypes.data)
        return l.reshape(self.height, self.width), n_lists

    def labels_refined(self):
        l = np.empty(self.n, np.uint32)
        lib().orc_get_labels_refined(self._h, l.ctypes.data)
        return l.reshape(self.height, self.width)

    def sat_exact(self) -> bool:
        return bool(lib().orc_sat_exact(self._h))

    def models(self):
        out = []
        for i in range(lib().orc_num_models(self._h)):
            coef = np.empty(4, np.float32)
            cen = np.empty(3, np.float32)
            cov = np.empty(9, np.float32)
            curv, label = C.c_float(), C.c_uint32()
            ns, nr, nc = C.c_int(), C.c_int(), C.c_int()
            lib().orc_get_model(self._h, i, coef.ctypes.data, cen.ctypes.data, cov.ctypes.data, C.byref(curv),
                                C.byref(label), C.byref(ns), C.byref(nr), C.byref(nc))
            inl = np.empty(nr.value, np.int32)
            con = np.empty(nc.value, np.int32)
            lib().orc_get_model_inliers(self._h, i, inl.ctypes.data)
            lib().orc_get_model_contour(self._h, i, con.ctypes.data)
            out.append(dict(coef=coef, centroid=cen, cov=cov.reshape(3, 3), curvature=curv.value, label=label.value,
                            n_segment=ns.value, inliers=inl, contour=con))
        return out

    def planes(self):
        out = []
        for i in range(lib().orc_num_planes(self._h)):
            coef = np.empty(4, np.float32)
            npts, nb, src = C.c_int(), C.c_int(), C.c_int()
            lib().orc_get_plane(self._h, i, coef.ctypes.data, C.byref(npts), C.byref(nb), C.byref(src))
            pts = np.empty(npts.value, POINT_DTYPE)
            bnd = np.empty(nb.value, POINT_DTYPE)
            lib().orc_get_plane_points(self._h, i, pts.ctypes.data)
            lib().orc_get_plane_boundary(self._h, i, bnd.ctypes.data)
            out.append(dict(coef=coef, points=pts, boundary=bnd, src=src.value))
        return out

    @property
    def n_real(self) -> int:
        return lib().orc_num_real_planes(self._h)

    @property
    def n_planes(self) -> int:
        return lib().orc_num_planes(self._h)

    def line_recs(self):
        n = lib().orc_num_line_recs(self._h)
        arr = (OrcLineRec * max(n, 1))()
        lib().orc_get_line_recs(self._h, arr)
        return [dict(plane=r.plane, round=r.round, n_points=r.n_points, iterations=r.iterations,
                     n_inliers=r.n_inliers, in_range=r.in_range, is_border=r.is_border, emitted=r.emitted,
                     coef=np.array(list(r.coef), np.float32)) for r in arr[:n]]

    def times(self):
        a, b = C.c_double(), C.c_double()
        lib().orc_get_times(self._h, C.byref(a), C.byref(b))
        return a.value, b.value


def run_batch(depth: np.ndarray, n_threads: int, cfg: OrcConfig | None = None):
    """Frame-parallel oracle over (n, rows, cols) float32 frames; returns (n_real[n], n_planes[n], t_plane_sum, t_splane_sum)."""
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    n, rows, cols = depth.shape
    cfg = cfg if cfg is not None else default_config()
    nr = np.zeros(n, np.int32)
    na = np.zeros(n, np.int32)
    a, b = C.c_double(), C.c_double()
    lib().orc_run_batch(C.byref(cfg), depth.ctypes.data, n, rows, cols, int(n_threads), nr.ctypes.data, na.ctypes.data,
                        C.byref(a), C.byref(b))
    return nr, na, a.value, b.value


def sat_exact_batch(depth: np.ndarray, n_threads: int, cfg: OrcConfig | None = None) -> np.ndarray:
    """Per frame: were all fp64 integral-image partial sums and window sums exact (error-free-transformation check)?"""
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    n, rows, cols = depth.shape
    cfg = cfg if cfg is not None else default_config()
    out = np.zeros(n, np.uint8)
    lib().orc_sat_exact_batch(C.byref(cfg), depth.ctypes.data, n, rows, cols, int(n_threads), out.ctypes.data)
    return out.astype(bool)


def chamfer(mask: np.ndarray, alt: int = 0) -> np.ndarray:
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w = mask.shape
    out = np.empty((h, w), np.float32)
    lib().orc_chamfer_alt(mask.ctypes.data, w, h, out.ctypes.data, alt)
    return out


def is_border_point(depth: np.ndarray, x: float, y: float, z: float, cfg: OrcConfig | None = None) -> bool:
    """Frame::IsBorderPoint (src/Frame.cc:1026-1056) of a camera-frame point against a CV_32F depth image."""
    cfg = cfg if cfg is not None else default_config()
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    return bool(lib().orc_is_border_point(C.byref(cfg), depth.ctypes.data, depth.shape[0], depth.shape[1],
                                          C.c_float(x), C.c_float(y), C.c_float(z)))


def eigen33_smallest(cov: np.ndarray):
    cov = np.ascontiguousarray(cov, dtype=np.float32).reshape(9)
    ev = C.c_float()
    vec = np.empty(3, np.float32)
    lib().orc_eigen33_smallest(cov.ctypes.data, C.byref(ev), vec.ctypes.data)
    return ev.value, vec


def eigen33_largest(cov: np.ndarray):
    cov = np.ascontiguousarray(cov, dtype=np.float32).reshape(9)
    evals = np.empty(3, np.float32)
    vec = np.empty(3, np.float32)
    lib().orc_eigen33_largest(cov.ctypes.data, evals.ctypes.data, vec.ctypes.data)
    return evals, vec


def sac_line(points: np.ndarray, threshold: float = float(np.float32(0.01)), max_iter: int = 1000, alt: int = 0):
    pts = np.ascontiguousarray(points, dtype=POINT_DTYPE)
    coef = np.empty(6, np.float32)
    inl = np.empty(max(len(pts), 1), np.int32)
    it = C.c_int()
    n = lib().orc_sac_line_alt(pts.ctypes.data, len(pts), float(threshold), int(max_iter), coef.ctypes.data,
                               inl.ctypes.data, C.byref(it), alt)
    return coef, inl[:n].copy(), it.value


def ransac_draws(n: int, n_draws: int) -> np.ndarray:
    out = np.empty((n_draws, 2), np.int32)
    lib().orc_ransac_draws(int(n), int(n_draws), out.ctypes.data)
    return out


def voxel_grid(points: np.ndarray, leaf):
    """pcl::VoxelGrid<PointXYZRGB> (PCL 1.8.0) on one cloud; returns (downsampled cloud, voxel index per output point)."""
    pts = np.ascontiguousarray(points, dtype=POINT_DTYPE)
    leaf3 = np.broadcast_to(np.asarray(leaf, np.float32), (3,)).copy()
    out = np.empty(max(len(pts), 1), POINT_DTYPE)
    idx = np.empty(max(len(pts), 1), np.int32)
    n = lib().orc_voxel_grid(pts.ctypes.data, len(pts), leaf3.ctypes.data, out.ctypes.data, idx.ctypes.data)
    return out[:n].copy(), idx[:n].copy()


def associate_planes(plane_w, map_w, boundaries, n_seen=None, dis_th=0.2, ang_th=0.8, ver_th=0.08716, par_th=0.9962):
    """Map::AssociatePlanesByBoundary (src/Map.cc:196-283) for one frame; returns (assoc, vertical, parallel, dist)."""
    plane_w = np.ascontiguousarray(plane_w, np.float32).reshape(-1, 4)
    map_w = np.ascontiguousarray(map_w, np.float32).reshape(-1, 4)
    n, n_map = len(plane_w), len(map_w)
    off = np.zeros(n_map + 1, np.int64)
    off[1:] = np.cumsum([len(b) for b in boundaries])
    pts = np.ascontiguousarray(np.concatenate([np.asarray(b, POINT_DTYPE) for b in boundaries]) if n_map else np.empty(0, POINT_DTYPE))
    a, v, p = (np.full(max(n, 1), -1, np.int32) for _ in range(3))
    d = np.zeros(max(n, 1), np.float32)
    lib().orc_associate_planes(plane_w.ctypes.data, n, map_w.ctypes.data, pts.ctypes.data, off.ctypes.data,
                               n_map if n_seen is None else n_seen, n_map, dis_th, ang_th, ver_th, par_th,
                               a.ctypes.data, v.ctypes.data, p.ctypes.data, d.ctypes.data)
    return a[:n], v[:n], p[:n], d[:n]


def transform_cloud(points: np.ndarray, m: np.ndarray) -> np.ndarray:
    """pcl::transformPointCloud(cloud, out, Matrix4d m) (MapPlane::UpdateBoundary, src/MapPlane.cc:144-147)"""
    pts = np.ascontiguousarray(points, dtype=POINT_DTYPE)
    m = np.ascontiguousarray(m, np.float64).reshape(4, 4)
    out = np.empty(len(pts), POINT_DTYPE)
    if len(pts):
        lib().orc_transform_cloud(pts.ctypes.data, len(pts), m.ctypes.data, out.ctypes.data)
    return out
