"""ctypes binding of the C ABI (include/spx.h, sp_slam_b200/libspx.so).

The library is sm_100a CUDA only.  Loading works anywhere (so the symbol table can be checked on a CPU box); every
compute entry point fails with SPX_ERR_CUDA when no B200-class device is present -- there is no CPU path.

``PlaneExtractor`` mirrors the two reference calls it replaces, Frame::ComputePlanesFromOrganizedPointCloud and
Frame::GeneratePlanesFromBoundries (/root/reference/src/Frame.cc:186,194), and hands back the same fields
(mvPlanePoints, mvBoundaryPoints, mvPlaneCoefficients, mnRealPlaneNum, mnPlaneNum; include/Frame.h:223-244).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

# The library runs a batch as frame groups on their own CUDA streams plus an upload and a download stream; streams beyond
# the driver's hardware-queue count (default 8) share a queue and serialise.  The variable is read when the CUDA context
# is created (spx_create sets it too, for processes in which the library is the first CUDA user).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspx.so")

SPX_OK, SPX_ERR_ARG, SPX_ERR_CUDA, SPX_ERR_STATE = 0, 1, 2, 3
SPX_FRAME_OVERFLOW = 1
SPX_FRAME_NONFINITE = 2
SPX_FRAME_SAT_UNPROVEN = 4
SPX_MAX_CAND, SPX_MAX_MODELS, SPX_MAX_PLANES, SPX_MAX_LINES = 96, 64, 128, 4

POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])
PLANE_DTYPE = np.dtype([("coef", "<f4", 4), ("n_points", "<i4"), ("n_boundary", "<i4"), ("points_off", "<i8"),
                        ("boundary_off", "<i8"), ("src", "<i4"), ("is_supposed", "<i4")])
HEADER_DTYPE = np.dtype([("n_real", "<i4"), ("n_planes", "<i4"), ("first_plane", "<i4"), ("flags", "<u4")])
MODEL_DTYPE = np.dtype([("coef", "<f4", 4), ("centroid", "<f4", 3), ("cov", "<f4", 9), ("curvature", "<f4"),
                        ("label", "<u4"), ("n_segment", "<i4"), ("n_inliers", "<i4"), ("n_contour", "<i4")])
EDGE_DTYPE = np.dtype([("kind", "<i4"), ("reserved", "<i4"), ("plane_w", "<f4", 4), ("measurement", "<f4", 4),
                       ("info", "<f8", 3), ("huber_delta", "<f8"), ("chi2_max", "<f8")])
LINE_DTYPE = np.dtype([("plane", "<i4"), ("round", "<i4"), ("n_points", "<i4"), ("iterations", "<i4"),
                       ("n_inliers", "<i4"), ("in_range", "<i4"), ("is_border", "<i4"), ("emitted", "<i4"),
                       ("coef", "<f4", 6)])

# every symbol include/spx.h declares (tests check the built library exports exactly these)
EXPORTS = (
    "spx_default_config", "spx_create", "spx_destroy", "spx_last_error", "spx_set_stream", "spx_extract",
    "spx_extract_batch", "spx_extract_batch_u16", "spx_extract_batch_device", "spx_fetch_results", "spx_fetch_planes",
    "spx_segment_from_normals", "spx_cloud_dims", "spx_get_times", "spx_last_launch_count", "spx_set_debug",
    "spx_set_profile", "spx_get_kernel_times", "spx_get_kernel_timeline", "spx_get_device_results",
    "spx_get_cloud", "spx_get_distance_map", "spx_get_normals", "spx_get_curvature", "spx_get_labels_raw", "spx_get_plane_ids",
    "spx_get_models", "spx_get_model_inliers", "spx_get_model_contour", "spx_get_lines",
    "spx_set_upload_mode", "spx_set_gather_threads", "spx_set_gather_share", "spx_host_gather_samples", "spx_host_register", "spx_host_unregister", "spx_get_transfer_bytes",
    "spx_get_group_timeline",
    "spx_extract_batch_compact", "spx_extract_batch_u16_compact", "spx_set_result_mode", "spx_fetch_compact",
    "spx_set_group_callback",
    "spx_voxel_grid", "spx_voxel_downsample_results", "spx_map_create", "spx_map_destroy", "spx_map_upload", "spx_map_associate",
    "spx_map_update_boundary", "spx_map_update_boundary_from_result", "spx_map_set_world_pos", "spx_map_get_boundary",
    "spx_pose_optimize_planes", "spx_plane_edge_errors",
)


class SpxConfig(C.Structure):
    _fields_ = [
        ("cloud_dis", C.c_int32), ("min_size", C.c_int32), ("angle_thr_deg", C.c_float), ("dist_thr", C.c_float),
        ("line_ratio", C.c_double), ("line_dist_thr", C.c_float),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
        ("max_depth_change_factor", C.c_float), ("normal_smoothing_size", C.c_float),
        ("ransac_max_iter", C.c_int32), ("enable_supposed", C.c_int32),
        ("max_frames", C.c_int32), ("max_rows", C.c_int32), ("max_cols", C.c_int32), ("device", C.c_int32),
        ("n_streams", C.c_int32), ("normal_method", C.c_int32),
    ]


class SpxBatchResult(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int32), ("n_planes_total", C.c_int32), ("n_points_total", C.c_int64),
        ("n_boundary_total", C.c_int64),
        ("frames", C.c_void_p), ("planes", C.c_void_p), ("points", C.c_void_p), ("boundary", C.c_void_p),
    ]


class SpxCompactResult(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int32), ("n_planes_total", C.c_int32), ("index_width", C.c_int32),
        ("cloud_width", C.c_int32), ("cloud_height", C.c_int32), ("cloud_dis", C.c_int32),
        ("n_index_total", C.c_int64), ("n_points_total", C.c_int64), ("n_boundary_total", C.c_int64),
        ("frames", C.c_void_p), ("planes", C.c_void_p), ("point_index", C.c_void_p), ("points", C.c_void_p),
        ("boundary", C.c_void_p),
    ]


GROUP_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.POINTER(SpxCompactResult))


class SpxDeviceResult(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int32), ("frames", C.c_void_p), ("planes", C.c_void_p), ("points", C.c_void_p),
        ("boundary", C.c_void_p), ("totals", C.c_void_p), ("planes_capacity", C.c_int64),
    ]


class SpxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"spx error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libspx.so; raises if the CUDA extension has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C sp_slam_b200/csrc` "
                               "(or `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
        L.spx_default_config.argtypes = [C.POINTER(SpxConfig)]
        L.spx_default_config.restype = None
        L.spx_create.argtypes = [C.POINTER(SpxConfig), C.POINTER(vp)]
        L.spx_destroy.argtypes = [vp]
        L.spx_destroy.restype = None
        L.spx_last_error.argtypes = [vp]
        L.spx_last_error.restype = C.c_char_p
        L.spx_set_stream.argtypes = [vp, vp]
        L.spx_set_debug.argtypes = [vp, i32]
        L.spx_extract.argtypes = [vp, vp, i32, i32, sz, C.POINTER(SpxBatchResult)]
        L.spx_extract_batch.argtypes = [vp, vp, i32, i32, i32, sz, sz, C.POINTER(SpxBatchResult)]
        L.spx_extract_batch_u16.argtypes = [vp, vp, i32, i32, i32, sz, sz, C.c_float, C.POINTER(SpxBatchResult)]
        L.spx_extract_batch_device.argtypes = [vp, vp, i32, i32, i32, sz, sz]
        L.spx_fetch_results.argtypes = [vp, C.POINTER(SpxBatchResult)]
        L.spx_fetch_planes.argtypes = [vp, C.POINTER(SpxBatchResult)]
        L.spx_segment_from_normals.argtypes = [vp, vp, i32, i32, sz, vp, C.POINTER(SpxBatchResult)]
        L.spx_cloud_dims.argtypes = [vp, i32, i32, C.POINTER(i32), C.POINTER(i32)]
        L.spx_get_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.spx_last_launch_count.argtypes = [vp]
        L.spx_get_device_results.argtypes = [vp, C.POINTER(SpxDeviceResult)]
        L.spx_set_profile.argtypes = [vp, i32]
        L.spx_get_kernel_times.argtypes = [vp, vp, vp, i32, C.POINTER(i32)]
        L.spx_get_kernel_timeline.argtypes = [vp, vp, vp, vp, i32, C.POINTER(i32)]
        L.spx_get_cloud.argtypes = [vp, i32, vp, vp, vp]
        L.spx_get_distance_map.argtypes = [vp, i32, vp]
        L.spx_get_normals.argtypes = [vp, i32, vp, vp, vp, vp]
        L.spx_get_curvature.argtypes = [vp, i32, vp]
        L.spx_get_labels_raw.argtypes = [vp, i32, vp, C.POINTER(i32)]
        L.spx_get_plane_ids.argtypes = [vp, i32, vp]
        L.spx_get_models.argtypes = [vp, i32, vp, C.POINTER(i32)]
        L.spx_get_model_inliers.argtypes = [vp, i32, i32, vp]
        L.spx_get_model_contour.argtypes = [vp, i32, i32, vp]
        L.spx_get_lines.argtypes = [vp, i32, vp, C.POINTER(i32)]
        L.spx_get_group_timeline.argtypes = [vp, vp, i32, C.POINTER(i32)]
        L.spx_voxel_grid.argtypes = [vp, vp, vp, i32, vp, vp, vp]
        L.spx_voxel_downsample_results.argtypes = [vp, C.c_float, i32]
        L.spx_map_create.argtypes = [vp, C.POINTER(vp)]
        L.spx_map_destroy.argtypes = [vp]
        L.spx_map_destroy.restype = None
        L.spx_map_upload.argtypes = [vp, vp, vp, vp, i32, i32]
        L.spx_map_associate.argtypes = [vp, vp, i32, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp]
        L.spx_map_update_boundary.argtypes = [vp, i32, vp, vp, i32]
        L.spx_map_update_boundary_from_result.argtypes = [vp, i32, vp, i32, i32, i32]
        L.spx_map_set_world_pos.argtypes = [vp, i32, vp]
        L.spx_map_get_boundary.argtypes = [vp, i32, vp, i32, C.POINTER(i32)]
        L.spx_pose_optimize_planes.argtypes = [vp, vp, i32, i32, i32, vp, vp, C.POINTER(i32)]
        L.spx_plane_edge_errors.argtypes = [vp, vp, i32, vp]
        L.spx_set_upload_mode.argtypes = [vp, i32]
        L.spx_set_gather_threads.argtypes = [vp, i32]
        L.spx_set_gather_share.argtypes = [vp, C.c_double]
        L.spx_host_gather_samples.argtypes = [vp, i32, i32, i32, sz, sz, i32, i32, i32, vp, sz]
        L.spx_host_register.argtypes = [vp, sz]
        L.spx_host_unregister.argtypes = [vp]
        u64p = C.POINTER(C.c_ulonglong)
        L.spx_get_transfer_bytes.argtypes = [vp, u64p, u64p, u64p]
        L.spx_extract_batch_compact.argtypes = [vp, vp, i32, i32, i32, sz, sz, C.POINTER(SpxCompactResult)]
        L.spx_extract_batch_u16_compact.argtypes = [vp, vp, i32, i32, i32, sz, sz, C.c_float, C.POINTER(SpxCompactResult)]
        L.spx_set_result_mode.argtypes = [vp, i32]
        L.spx_fetch_compact.argtypes = [vp, C.POINTER(SpxCompactResult)]
        L.spx_set_group_callback.argtypes = [vp, GROUP_FN, vp]
        for name in EXPORTS:
            getattr(L, name)   # AttributeError here = the library is stale
        _lib = L
    return _lib


def host_register(arr: np.ndarray):
    """Page-lock a host array in place (spx_host_register); pair with host_unregister."""
    rc = lib().spx_host_register(arr.ctypes.data, arr.nbytes)
    if rc != SPX_OK:
        raise SpxError(rc, (lib().spx_last_error(None) or b"").decode())


def gather_samples(depth: np.ndarray, cloud_dis: int = 5, n_groups: int = 1, n_threads: int = 2, row_floats: int | None = None) -> np.ndarray:
    """spx_host_gather_samples: the organized cloud's samples depth[f, ::dis, ::dis] (src/Frame.cc:857-872) of a float batch
    (frames, rows, cols; the last axis contiguous), staged as the gathered upload stages them.  Host code only."""
    if depth.ndim == 2:
        depth = depth[None]
    assert depth.dtype == np.float32 and depth.ndim == 3 and depth.strides[2] == 4
    n, rows, cols = depth.shape
    h, w = -(-rows // cloud_dis), -(-cols // cloud_dis)
    rf = (w + 3) & ~3 if row_floats is None else row_floats
    out = np.full((n, h, rf), np.nan, np.float32)
    rc = lib().spx_host_gather_samples(depth.ctypes.data, n, rows, cols, depth.strides[1], depth.strides[0] if n > 1 else depth.strides[1] * rows,
                                       cloud_dis, n_groups, n_threads, out.ctypes.data, rf)
    if rc != SPX_OK:
        raise SpxError(rc, "spx_host_gather_samples")
    return out


def host_unregister(arr: np.ndarray):
    rc = lib().spx_host_unregister(arr.ctypes.data)
    if rc != SPX_OK:
        raise SpxError(rc, (lib().spx_last_error(None) or b"").decode())


def default_config(**overrides) -> SpxConfig:
    cfg = SpxConfig()
    lib().spx_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def _view(ptr, count, dtype):
    if not ptr or count == 0:
        return np.empty(0, dtype)
    buf = (C.c_char * (int(count) * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(count))


@dataclass
class FramePlanes:
    """The plane fields of one Frame (include/Frame.h:223-244)."""
    mnRealPlaneNum: int
    mnPlaneNum: int
    mvPlaneCoefficients: np.ndarray          # (mnPlaneNum, 4) float32, unit normal + d, d >= 0
    mvPlanePoints: list                      # per plane: POINT_DTYPE array
    mvBoundaryPoints: list                   # per plane: POINT_DTYPE array
    src: np.ndarray                          # real: segmentation model index; supposed: parent plane index
    is_supposed: np.ndarray
    flags: int


class BatchResult:
    """Copies of the context-owned result buffers of one extract call."""

    def __init__(self, r: SpxBatchResult, copy: bool = True):
        f = _view(r.frames, r.n_frames, HEADER_DTYPE)
        p = _view(r.planes, r.n_planes_total, PLANE_DTYPE)
        pts = _view(r.points, r.n_points_total, POINT_DTYPE)
        bnd = _view(r.boundary, r.n_boundary_total, POINT_DTYPE)
        if copy:
            f, p, pts, bnd = f.copy(), p.copy(), pts.copy(), bnd.copy()
        self.frames, self.planes, self.points, self.boundary = f, p, pts, bnd

    def __len__(self):
        return len(self.frames)

    def frame(self, i: int) -> FramePlanes:
        h = self.frames[i]
        pl = self.planes[h["first_plane"]: h["first_plane"] + h["n_planes"]]
        have_clouds = len(self.points) > 0 or len(self.boundary) > 0
        pts = [self.points[q["points_off"]: q["points_off"] + q["n_points"]] for q in pl] if have_clouds else []
        bnd = [self.boundary[q["boundary_off"]: q["boundary_off"] + q["n_boundary"]] for q in pl] if have_clouds else []
        return FramePlanes(int(h["n_real"]), int(h["n_planes"]), pl["coef"].copy(), pts, bnd, pl["src"].copy(),
                           pl["is_supposed"].copy(), int(h["flags"]))


RGBA_CLOUD = 0xFF0000FA      # a = 255, (r, g, b) = (0, 0, 250): the colour of every organized-cloud point (src/Frame.cc:866-868)


def backproject(idx: np.ndarray, depth: np.ndarray, cfg, w: int, dis: int) -> np.ndarray:
    """inputCloud.points[idx] (src/Frame.cc:857-870) in fp32, one rounding per operation: what the host adapter rebuilds from
    a compact result.  depth: one CV_32F frame."""
    idx = idx.astype(np.int64)
    r, c = idx // w, idx % w
    z = depth[r * dis, c * dis].astype(np.float32)
    out = np.empty(len(idx), POINT_DTYPE)
    out["x"] = ((c * dis).astype(np.float32) - np.float32(cfg.cx)) * z / np.float32(cfg.fx)
    out["y"] = ((r * dis).astype(np.float32) - np.float32(cfg.cy)) * z / np.float32(cfg.fy)
    out["z"] = z
    out["rgba"] = RGBA_CLOUD
    return out


class CompactResult:
    """spx_compact_result: the real planes' clouds as ordered inlier index lists (2 or 4 bytes per point)."""

    def __init__(self, r: SpxCompactResult, copy: bool = True):
        self.index_width = int(r.index_width)
        self.cloud_width, self.cloud_height, self.cloud_dis = int(r.cloud_width), int(r.cloud_height), int(r.cloud_dis)
        f = _view(r.frames, r.n_frames, HEADER_DTYPE)
        p = _view(r.planes, r.n_planes_total, PLANE_DTYPE)
        ix = _view(r.point_index, r.n_index_total, np.dtype("<u2" if r.index_width == 2 else "<u4"))
        pts = _view(r.points, r.n_points_total, POINT_DTYPE)
        bnd = _view(r.boundary, r.n_boundary_total, POINT_DTYPE)
        if copy:
            f, p, ix, pts, bnd = f.copy(), p.copy(), ix.copy(), pts.copy(), bnd.copy()
        self.frames, self.planes, self.point_index, self.points, self.boundary = f, p, ix, pts, bnd

    def __len__(self):
        return len(self.frames)

    @property
    def nbytes(self):
        return self.frames.nbytes + self.planes.nbytes + self.point_index.nbytes + self.points.nbytes + self.boundary.nbytes

    def frame(self, i: int, depth: np.ndarray, cfg) -> FramePlanes:
        """The Frame fields of frame i, the real planes' clouds rebuilt from `depth` (that frame's CV_32F image)."""
        h = self.frames[i]
        pl = self.planes[h["first_plane"]: h["first_plane"] + h["n_planes"]]
        pts = []
        for q in pl:
            if q["is_supposed"]:
                pts.append(self.points[q["points_off"]: q["points_off"] + q["n_points"]])
            else:
                pts.append(backproject(self.point_index[q["points_off"]: q["points_off"] + q["n_points"]], depth, cfg,
                                       self.cloud_width, self.cloud_dis))
        bnd = [self.boundary[q["boundary_off"]: q["boundary_off"] + q["n_boundary"]] for q in pl]
        return FramePlanes(int(h["n_real"]), int(h["n_planes"]), pl["coef"].copy(), pts, bnd, pl["src"].copy(),
                           pl["is_supposed"].copy(), int(h["flags"]))


class PlaneExtractor:
    """One context = one CUDA device + fixed capacity (frames per batch, image size)."""

    def __init__(self, cfg: SpxConfig | None = None, debug: bool = False, **overrides):
        self.cfg = cfg if cfg is not None else default_config(**overrides)
        self._h = C.c_void_p()
        rc = lib().spx_create(C.byref(self.cfg), C.byref(self._h))
        if rc != SPX_OK:
            self._h = C.c_void_p()
            raise SpxError(rc, (lib().spx_last_error(None) or b"").decode())
        if debug:
            self._ck(lib().spx_set_debug(self._h, 1))

    def close(self):
        if getattr(self, "_h", None):
            lib().spx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != SPX_OK:
            raise SpxError(rc, (lib().spx_last_error(self._h) or b"").decode())

    # ---- extraction ----
    def extract(self, depth: np.ndarray) -> FramePlanes:
        """One frame: CV_32F metres, rows x cols (host)."""
        return self.extract_batch(depth[None]).frame(0)

    def extract_batch(self, depth: np.ndarray) -> BatchResult:
        """(n_frames, rows, cols) float32 host array; H2D copy, kernels and D2H copy inside the call."""
        if depth.dtype != np.float32 or depth.ndim != 3 or depth.strides[2] != 4:
            depth = np.ascontiguousarray(depth, dtype=np.float32)
        n, rows, cols = depth.shape
        r = SpxBatchResult()
        self._ck(lib().spx_extract_batch(self._h, depth.ctypes.data, n, rows, cols, depth.strides[1],
                                         depth.strides[0], C.byref(r)))
        return BatchResult(r)

    def extract_batch_ptr(self, host_ptr: int, n: int, rows: int, cols: int, copy: bool = False) -> BatchResult:
        """Same, from a raw host pointer (e.g. a pinned torch tensor's data_ptr()); tight row-major layout."""
        r = SpxBatchResult()
        self._ck(lib().spx_extract_batch(self._h, host_ptr, n, rows, cols, cols * 4, rows * cols * 4, C.byref(r)))
        return BatchResult(r, copy=copy)

    def extract_batch_u16(self, depth_u16: np.ndarray, depth_map_factor: float = float(np.float32(1.0) / np.float32(5000.0)),
                          copy: bool = True) -> BatchResult:
        """(n_frames, rows, cols) uint16 host array (e.g. TUM PNG depth); depth = float(d) * depth_map_factor on the device
        (Tracking::GrabImageRGBD's convertTo, src/Tracking.cc:230-231)."""
        if depth_u16.dtype != np.uint16 or depth_u16.ndim != 3 or depth_u16.strides[2] != 2:
            depth_u16 = np.ascontiguousarray(depth_u16, dtype=np.uint16)
        n, rows, cols = depth_u16.shape
        r = SpxBatchResult()
        self._ck(lib().spx_extract_batch_u16(self._h, depth_u16.ctypes.data, n, rows, cols, depth_u16.strides[1],
                                             depth_u16.strides[0], C.c_float(depth_map_factor), C.byref(r)))
        return BatchResult(r, copy=copy)

    def extract_batch_u16_ptr(self, host_ptr: int, n: int, rows: int, cols: int, depth_map_factor: float, copy: bool = False):
        r = SpxBatchResult()
        self._ck(lib().spx_extract_batch_u16(self._h, host_ptr, n, rows, cols, cols * 2, rows * cols * 2,
                                             C.c_float(depth_map_factor), C.byref(r)))
        return BatchResult(r, copy=copy)

    # ---- compact results ----
    def extract_batch_compact(self, depth: np.ndarray, copy: bool = True) -> CompactResult:
        if depth.dtype != np.float32 or depth.ndim != 3 or depth.strides[2] != 4:
            depth = np.ascontiguousarray(depth, dtype=np.float32)
        n, rows, cols = depth.shape
        r = SpxCompactResult()
        self._ck(lib().spx_extract_batch_compact(self._h, depth.ctypes.data, n, rows, cols, depth.strides[1], depth.strides[0], C.byref(r)))
        return CompactResult(r, copy=copy)

    def extract_batch_compact_ptr(self, host_ptr: int, n: int, rows: int, cols: int, copy: bool = False) -> CompactResult:
        r = SpxCompactResult()
        self._ck(lib().spx_extract_batch_compact(self._h, host_ptr, n, rows, cols, cols * 4, rows * cols * 4, C.byref(r)))
        return CompactResult(r, copy=copy)

    def extract_batch_u16_compact(self, depth_u16: np.ndarray, depth_map_factor: float = float(np.float32(1.0) / np.float32(5000.0)),
                                  copy: bool = True) -> CompactResult:
        if depth_u16.dtype != np.uint16 or depth_u16.ndim != 3 or depth_u16.strides[2] != 2:
            depth_u16 = np.ascontiguousarray(depth_u16, dtype=np.uint16)
        n, rows, cols = depth_u16.shape
        r = SpxCompactResult()
        self._ck(lib().spx_extract_batch_u16_compact(self._h, depth_u16.ctypes.data, n, rows, cols, depth_u16.strides[1],
                                                     depth_u16.strides[0], C.c_float(depth_map_factor), C.byref(r)))
        return CompactResult(r, copy=copy)

    def extract_batch_u16_compact_ptr(self, host_ptr: int, n: int, rows: int, cols: int, depth_map_factor: float, copy: bool = False):
        r = SpxCompactResult()
        self._ck(lib().spx_extract_batch_u16_compact(self._h, host_ptr, n, rows, cols, cols * 2, rows * cols * 2,
                                                     C.c_float(depth_map_factor), C.byref(r)))
        return CompactResult(r, copy=copy)

    def set_result_mode(self, compact: bool):
        """What extract_device packs: point clouds (fetch) or compact results (fetch_compact)."""
        self._ck(lib().spx_set_result_mode(self._h, 1 if compact else 0))

    def fetch_compact(self, copy: bool = True) -> CompactResult:
        r = SpxCompactResult()
        self._ck(lib().spx_fetch_compact(self._h, C.byref(r)))
        return CompactResult(r, copy=copy)

    def set_group_callback(self, fn):
        """fn(frame0, frame1, CompactResult view) as soon as a frame group of a compact host-input call is on the host; None = off."""
        if fn is None:
            self._cb = None
            self._ck(lib().spx_set_group_callback(self._h, GROUP_FN(), None))   # a NULL function pointer
            return
        self._cb = GROUP_FN(lambda user, f0, f1, view: fn(f0, f1, CompactResult(view.contents, copy=True)))
        self._ck(lib().spx_set_group_callback(self._h, self._cb, None))

    def extract_device(self, dev_ptr: int, n: int, rows: int, cols: int, pitch: int | None = None,
                       frame_stride: int | None = None):
        """Asynchronous, device-resident depth; results stay on the device until fetch()."""
        pitch = cols * 4 if pitch is None else pitch
        frame_stride = pitch * rows if frame_stride is None else frame_stride
        self._ck(lib().spx_extract_batch_device(self._h, dev_ptr, n, rows, cols, pitch, frame_stride))

    def fetch(self, clouds: bool = True, copy: bool = True) -> BatchResult:
        r = SpxBatchResult()
        self._ck((lib().spx_fetch_results if clouds else lib().spx_fetch_planes)(self._h, C.byref(r)))
        return BatchResult(r, copy=copy)

    def segment_from_normals(self, depth: np.ndarray, normals: np.ndarray) -> FramePlanes:
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        normals = np.ascontiguousarray(normals, dtype=np.float32)
        rows, cols = depth.shape
        r = SpxBatchResult()
        self._ck(lib().spx_segment_from_normals(self._h, depth.ctypes.data, rows, cols, cols * 4, normals.ctypes.data,
                                                C.byref(r)))
        return BatchResult(r).frame(0)

    def device_results(self) -> SpxDeviceResult:
        r = SpxDeviceResult()
        self._ck(lib().spx_get_device_results(self._h, C.byref(r)))
        return r

    def set_stream(self, cuda_stream: int | None):
        self._ck(lib().spx_set_stream(self._h, cuda_stream))
        self._stream = cuda_stream or None

    @property
    def stream(self):
        """The caller-owned cudaStream_t the context is bound to (None: the context's own non-blocking stream)."""
        return getattr(self, "_stream", None)

    def cloud_dims(self, rows: int, cols: int):
        w, h = C.c_int(), C.c_int()
        self._ck(lib().spx_cloud_dims(self._h, rows, cols, C.byref(w), C.byref(h)))
        return w.value, h.value

    def times(self):
        a, b = C.c_double(), C.c_double()
        self._ck(lib().spx_get_times(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def launches(self) -> int:
        return lib().spx_last_launch_count(self._h)

    def set_upload_mode(self, mode: int):
        """0 automatic, 1 whole image, 2 sparse (sampled rows uploaded, border-test window sectors fetched on demand) when page-locked,
        3 gathered (host threads stage the organized cloud's samples, only those are uploaded) when page-locked and float."""
        self._ck(lib().spx_set_upload_mode(self._h, mode))

    def set_gather_threads(self, n: int):
        """host threads of the gathered route (0 = half the hardware threads, at most 16)"""
        self._ck(lib().spx_set_gather_threads(self._h, n))

    def set_gather_share(self, share: float):
        """automatic upload mode: share of a batch (its last frame groups) that takes the gathered route"""
        self._ck(lib().spx_set_gather_share(self._h, float(share)))

    def transfer_bytes(self):
        """(bytes uploaded by copies, bytes fetched from the caller's pinned image by the device, bytes copied back) of the last
        host-input extract."""
        a, b, d = C.c_ulonglong(), C.c_ulonglong(), C.c_ulonglong()
        self._ck(lib().spx_get_transfer_bytes(self._h, C.byref(a), C.byref(b), C.byref(d)))
        return a.value, b.value, d.value

    def group_timeline(self) -> np.ndarray:
        """(groups, 7) ms after the start of the last host-input call: started, depth on the device, real planes final,
        last kernel ended, results on the host (device clock); enqueue began, totals seen (host clock)."""
        t = np.zeros((64, 7), np.float32)
        n = C.c_int()
        self._ck(lib().spx_get_group_timeline(self._h, t.ctypes.data, 64, C.byref(n)))
        return t[:n.value].copy()

    # ---- SURVEY 8(f) rows ----
    def voxel_grid(self, clouds, leaf):
        """pcl::VoxelGrid<PointXYZRGB> with leaf size `leaf` (scalar or 3 values) on a list of POINT_DTYPE arrays (host);
        returns the list of downsampled clouds (PCL's order: ascending voxel index)."""
        leaf3 = np.broadcast_to(np.asarray(leaf, np.float32), (3,)).copy()
        off = np.zeros(len(clouds) + 1, np.int64)
        off[1:] = np.cumsum([len(c) for c in clouds])
        pts = np.ascontiguousarray(np.concatenate([np.asarray(c, POINT_DTYPE) for c in clouds]) if len(clouds) else np.empty(0, POINT_DTYPE))
        out = np.empty(max(len(pts), 1), POINT_DTYPE)
        out_off = np.zeros(len(clouds) + 1, np.int64)
        self._ck(lib().spx_voxel_grid(self._h, pts.ctypes.data, off.ctypes.data, len(clouds), leaf3.ctypes.data, out.ctypes.data,
                                      out_off.ctypes.data))
        return [out[out_off[s]:out_off[s + 1]].copy() for s in range(len(clouds))]

    def voxel_downsample_results(self, leaf: float, which: int):
        """Downsample, on the device, every plane's cloud (which=0) or contour (which=1) of the last extract_device call."""
        self._ck(lib().spx_voxel_downsample_results(self._h, C.c_float(leaf), which))

    def set_profile(self, on: bool):
        self._ck(lib().spx_set_profile(self._h, 1 if on else 0))

    def kernel_times(self):
        """[(kernel name, device ms)] of the last extract call (needs set_profile(True) before it)."""
        cap = 1024
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        n = C.c_int()
        self._ck(lib().spx_get_kernel_times(self._h, names, ms, cap, C.byref(n)))
        return [(names[k].decode(), float(ms[k])) for k in range(min(n.value, cap))]

    def kernel_timeline(self):
        """[(kernel name, start ms, end ms)] of the last extract call, relative to its beginning."""
        cap = 1024
        names = (C.c_char_p * cap)()
        t0 = (C.c_float * cap)()
        t1 = (C.c_float * cap)()
        n = C.c_int()
        self._ck(lib().spx_get_kernel_timeline(self._h, names, t0, t1, cap, C.byref(n)))
        return [(names[k].decode(), float(t0[k]), float(t1[k])) for k in range(min(n.value, cap))]

    # ---- debug taps (need debug=True) ----
    def _n(self, rows, cols):
        w, h = self.cloud_dims(rows, cols)
        return w, h, w * h

    def cloud(self, frame, n):
        a = [np.empty(n, np.float32) for _ in range(3)]
        self._ck(lib().spx_get_cloud(self._h, frame, *(x.ctypes.data for x in a)))
        return np.stack(a)

    def distance_map(self, frame, n):
        d = np.empty(n, np.float32)
        self._ck(lib().spx_get_distance_map(self._h, frame, d.ctypes.data))
        return d

    def normals(self, frame, n):
        a = [np.empty(n, np.float32) for _ in range(4)]
        self._ck(lib().spx_get_normals(self._h, frame, *(x.ctypes.data for x in a)))
        return np.stack(a[:3]), a[3]

    def curvature(self, frame, n):
        d = np.empty(n, np.float32)
        self._ck(lib().spx_get_curvature(self._h, frame, d.ctypes.data))
        return d

    def labels_raw(self, frame, n):
        l = np.empty(n, np.uint32)
        k = C.c_int()
        self._ck(lib().spx_get_labels_raw(self._h, frame, l.ctypes.data, C.byref(k)))
        return l, k.value

    def plane_ids(self, frame, n):
        p = np.empty(n, np.int8)
        self._ck(lib().spx_get_plane_ids(self._h, frame, p.ctypes.data))
        return p

    def models(self, frame):
        m = np.zeros(SPX_MAX_MODELS, MODEL_DTYPE)
        k = C.c_int()
        self._ck(lib().spx_get_models(self._h, frame, m.ctypes.data, C.byref(k)))
        out = []
        for i in range(k.value):
            inl = np.full(int(m[i]["n_inliers"]), -1, np.int32)
            con = np.empty(int(m[i]["n_contour"]), np.int32)
            if len(inl):
                self._ck(lib().spx_get_model_inliers(self._h, frame, i, inl.ctypes.data))
            if len(con):
                self._ck(lib().spx_get_model_contour(self._h, frame, i, con.ctypes.data))
            out.append(dict(coef=m[i]["coef"].copy(), centroid=m[i]["centroid"].copy(),
                            cov=m[i]["cov"].reshape(3, 3).copy(), curvature=float(m[i]["curvature"]),
                            label=int(m[i]["label"]), n_segment=int(m[i]["n_segment"]), inliers=inl, contour=con))
        return out

    def lines(self, frame):
        l = np.zeros(SPX_MAX_MODELS * SPX_MAX_LINES, LINE_DTYPE)
        k = C.c_int()
        self._ck(lib().spx_get_lines(self._h, frame, l.ctypes.data, C.byref(k)))
        return l[:k.value].copy()


HOST_LIB_PATH = os.path.join(_HERE, "libspx_host.so")
POINT32_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("data_w", "<f4"), ("rgba", "<u4"), ("pad", "<u4", 3)])   # pcl::PointXYZRGB
_host_lib = None


def host_lib():
    """libspx_host.so: the C++ host adapter (sp_slam_b200/host/SequencePlanes.h) behind a plain-C handle."""
    global _host_lib
    if _host_lib is None:
        lib()
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} is missing: build it with `make -C sp_slam_b200/csrc`")
        H = C.CDLL(HOST_LIB_PATH)
        vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
        H.spx_seq_create.argtypes = [C.POINTER(SpxConfig), i32]
        H.spx_seq_create.restype = vp
        H.spx_seq_destroy.argtypes = [vp]
        H.spx_seq_destroy.restype = None
        H.spx_seq_threads.argtypes = [vp]
        H.spx_seq_context.argtypes = [vp]
        H.spx_seq_context.restype = vp
        H.spx_seq_last_error.restype = C.c_char_p
        H.spx_seq_process.argtypes = [vp, vp, i32, i32, i32, sz, sz]
        H.spx_seq_process.restype = C.c_double
        H.spx_seq_process_clouds.argtypes = [vp, vp, i32, i32, i32, sz, sz]
        H.spx_seq_process_clouds.restype = C.c_double
        H.spx_seq_process_u16.argtypes = [vp, vp, i32, i32, i32, sz, sz, C.c_float]
        H.spx_seq_process_u16.restype = C.c_double
        H.spx_seq_summary.argtypes = [vp, vp]
        H.spx_seq_summary.restype = None
        H.spx_seq_frame.argtypes = [vp, i32, vp]
        H.spx_seq_plane.argtypes = [vp, i32, i32, vp, vp, C.POINTER(vp), C.POINTER(vp)]
        H.spx_seq_hash.argtypes = [vp]
        H.spx_seq_hash.restype = C.c_ulonglong
        _host_lib = H
    return _host_lib


class SequenceAdapter:
    """spx_host::SequencePlanes: a whole depth sequence through the C++ host adapter, every Frame field of every frame filled
    (32-byte pcl::PointXYZRGB-layout clouds) by a pool of host threads while later frame groups are still on the device."""

    def __init__(self, cfg: SpxConfig, n_threads: int = 0):
        self.cfg = cfg
        self._s = host_lib().spx_seq_create(C.byref(cfg), n_threads)
        if not self._s:
            raise SpxError(SPX_ERR_CUDA, (host_lib().spx_seq_last_error() or b"").decode())
        self.threads = host_lib().spx_seq_threads(self._s)

    def close(self):
        if getattr(self, "_s", None):
            host_lib().spx_seq_destroy(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process_ptr(self, host_ptr: int, n: int, rows: int, cols: int) -> float:
        """One batch of tight CV_32F frames; returns the wall time of the call in ms."""
        ms = host_lib().spx_seq_process(self._s, host_ptr, n, rows, cols, cols * 4, rows * cols * 4)
        if ms < 0:
            raise SpxError(SPX_ERR_CUDA, (host_lib().spx_seq_last_error() or b"").decode())
        return ms

    def process_clouds_ptr(self, host_ptr: int, n: int, rows: int, cols: int) -> float:
        """The same with the real planes' clouds crossing PCIe as 16-byte points (spx_extract_batch) and only widened on the host."""
        ms = host_lib().spx_seq_process_clouds(self._s, host_ptr, n, rows, cols, cols * 4, rows * cols * 4)
        if ms < 0:
            raise SpxError(SPX_ERR_CUDA, (host_lib().spx_seq_last_error() or b"").decode())
        return ms

    def process_u16_ptr(self, host_ptr: int, n: int, rows: int, cols: int, factor: float) -> float:
        ms = host_lib().spx_seq_process_u16(self._s, host_ptr, n, rows, cols, cols * 2, rows * cols * 2, C.c_float(factor))
        if ms < 0:
            raise SpxError(SPX_ERR_CUDA, (host_lib().spx_seq_last_error() or b"").decode())
        return ms

    def summary(self):
        """(planes, points of mvPlanePoints, points of mvBoundaryPoints, bytes of the filled fields) of the last batch."""
        out = (C.c_longlong * 4)()
        host_lib().spx_seq_summary(self._s, out)
        return tuple(int(v) for v in out)

    def hash(self) -> int:
        return int(host_lib().spx_seq_hash(self._s))

    def frame(self, f: int) -> FramePlanes:
        """Copies of the fields of frame f (clouds as POINT32_DTYPE arrays) plus (width, height) of every cloud in `dims`."""
        hd = (C.c_int * 3)()
        if host_lib().spx_seq_frame(self._s, f, hd):
            raise IndexError(f)
        coefs, pts, bnd, dims = [], [], [], []
        for i in range(hd[1]):
            coef = (C.c_float * 4)()
            sizes = (C.c_int * 6)()
            p, b = C.c_void_p(), C.c_void_p()
            host_lib().spx_seq_plane(self._s, f, i, coef, sizes, C.byref(p), C.byref(b))
            coefs.append(np.array(coef[:], np.float32))
            pts.append(_view(p.value, sizes[0], POINT32_DTYPE).copy())
            bnd.append(_view(b.value, sizes[3], POINT32_DTYPE).copy())
            dims.append(((sizes[1], sizes[2]), (sizes[4], sizes[5])))
        fp = FramePlanes(hd[0], hd[1], np.array(coefs, np.float32).reshape(-1, 4), pts, bnd, np.zeros(hd[1], np.int32),
                         np.zeros(hd[1], np.int32), hd[2])
        fp.dims = dims
        return fp


class PlaneMap:
    """Device-resident mirror of the map planes for Map::AssociatePlanesByBoundary (src/Map.cc:196-283)."""

    def __init__(self, ext: PlaneExtractor):
        self._ext = ext
        self._m = C.c_void_p()
        ext._ck(lib().spx_map_create(ext._h, C.byref(self._m)))

    def close(self):
        if getattr(self, "_m", None):
            lib().spx_map_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, map_w: np.ndarray, boundaries: list, n_seen: int | None = None):
        """map_w: (n_map, 4) world coefficients in the reference's visiting order; boundaries: per map plane a POINT_DTYPE
        cloud (world frame); the first n_seen planes are mspMapPlanes, the rest mspNotSeenMapPlanes."""
        map_w = np.ascontiguousarray(map_w, np.float32).reshape(-1, 4)
        n_map = len(map_w)
        off = np.zeros(n_map + 1, np.int64)
        off[1:] = np.cumsum([len(b) for b in boundaries])
        pts = np.ascontiguousarray(np.concatenate([np.asarray(b, POINT_DTYPE) for b in boundaries]) if n_map else np.empty(0, POINT_DTYPE))
        self._ext._ck(lib().spx_map_upload(self._m, map_w.ctypes.data, pts.ctypes.data, off.ctypes.data,
                                           n_map if n_seen is None else n_seen, n_map))

    def associate(self, plane_w: np.ndarray, dis_th=0.2, ang_th=0.8, ver_th=0.08716, par_th=0.9962):
        """Returns (assoc, vertical, parallel, dist): per frame plane the index of the matched map plane or -1
        (defaults: Examples/RGB-D/TUM1.yaml:83-88)."""
        plane_w = np.ascontiguousarray(plane_w, np.float32).reshape(-1, 4)
        n = len(plane_w)
        a, v, p = (np.full(max(n, 1), -1, np.int32) for _ in range(3))
        d = np.zeros(max(n, 1), np.float32)
        self._ext._ck(lib().spx_map_associate(self._m, plane_w.ctypes.data, n, C.c_float(dis_th), C.c_float(ang_th), C.c_float(ver_th),
                                              C.c_float(par_th), a.ctypes.data, v.ctypes.data, p.ctypes.data, d.ctypes.data))
        return a[:n], v[:n], p[:n], d[:n]

    def update_boundary(self, j: int, transform: np.ndarray, cloud: np.ndarray):
        """MapPlane::UpdateBoundary: map plane j's boundary becomes transform (4x4 double) applied to `cloud` (host)."""
        t = np.ascontiguousarray(transform, np.float64).reshape(4, 4)
        pts = np.ascontiguousarray(cloud, POINT_DTYPE)
        self._ext._ck(lib().spx_map_update_boundary(self._m, j, t.ctypes.data, pts.ctypes.data, len(pts)))

    def update_boundary_from_result(self, j: int, transform: np.ndarray, frame: int, plane: int, n_boundary: int):
        """The same with the boundary of plane `plane` of frame `frame` of the last extract, read on the device."""
        t = np.ascontiguousarray(transform, np.float64).reshape(4, 4)
        self._ext._ck(lib().spx_map_update_boundary_from_result(self._m, j, t.ctypes.data, frame, plane, n_boundary))

    def set_world_pos(self, j: int, coef_w):
        w = np.ascontiguousarray(coef_w, np.float32).reshape(4)
        self._ext._ck(lib().spx_map_set_world_pos(self._m, j, w.ctypes.data))

    def boundary(self, j: int) -> np.ndarray:
        n = C.c_int()
        self._ext._ck(lib().spx_map_get_boundary(self._m, j, None, 0, C.byref(n)))
        out = np.empty(max(n.value, 1), POINT_DTYPE)
        self._ext._ck(lib().spx_map_get_boundary(self._m, j, out.ctypes.data, n.value, C.byref(n)))
        return out[:n.value].copy()


def plane_edges(kinds, planes_w, measurements, angle_info=1.0, distance_info=100.0, parallel_info=0.5, vertical_info=0.5,
                chi=300.0, vp_chi=300.0, not_seen=None) -> np.ndarray:
    """The edge records Optimizer::PoseOptimization builds (src/Optimizer.cc:695-907) from the YAML keys Plane.AngleInfo,
    DistanceInfo, ParallelInfo, VerticalInfo, Chi, VPChi (defaults: Examples/RGB-D/TUM1.yaml:89-94)."""
    kinds = np.asarray(kinds, np.int32)
    e = np.zeros(len(kinds), EDGE_DTYPE)
    e["kind"] = kinds
    e["plane_w"] = np.asarray(planes_w, np.float32).reshape(-1, 4)
    e["measurement"] = np.asarray(measurements, np.float32).reshape(-1, 4)
    a, d = 3282.8 / (angle_info * angle_info), distance_info * distance_info
    par, ver = 3282.8 / (parallel_info * parallel_info), 3282.8 / (vertical_info * vertical_info)
    for i, k in enumerate(kinds):
        if k == 0:
            f = 2.0 if (not_seen is not None and not_seen[i]) else 1.0
            e["info"][i] = (f * a, f * a, f * d)
            e["huber_delta"][i], e["chi2_max"][i] = np.float32(np.sqrt(chi)), chi
        else:
            w = par if k == 1 else ver
            e["info"][i] = (w, w, 0.0)
            e["huber_delta"][i], e["chi2_max"][i] = np.float32(np.sqrt(vp_chi)), vp_chi
    return e


def pose_optimize_planes(Tcw: np.ndarray, edges: np.ndarray, rounds: int = 4, iterations: int = 10):
    """The plane part of Optimizer::PoseOptimization (host code, no GPU): returns (Tcw optimised, outlier flags, chi2, nBad)."""
    T = np.ascontiguousarray(Tcw, np.float64).reshape(4, 4).copy()
    edges = np.ascontiguousarray(edges, EDGE_DTYPE)
    out = np.zeros(max(len(edges), 1), np.uint8)
    chi2 = np.zeros(max(len(edges), 1), np.float64)
    bad = C.c_int()
    rc = lib().spx_pose_optimize_planes(T.ctypes.data, edges.ctypes.data, len(edges), rounds, iterations, out.ctypes.data,
                                        chi2.ctypes.data, C.byref(bad))
    if rc != SPX_OK:
        raise SpxError(rc, "spx_pose_optimize_planes: bad argument")
    return T, out[:len(edges)].astype(bool), chi2[:len(edges)], bad.value


def plane_edge_errors(Tcw: np.ndarray, edges: np.ndarray) -> np.ndarray:
    """computeError of every edge at the pose Tcw: (n, 3) doubles (third column 0 for the 2-d edges)."""
    T = np.ascontiguousarray(Tcw, np.float64).reshape(4, 4)
    edges = np.ascontiguousarray(edges, EDGE_DTYPE)
    err = np.zeros((max(len(edges), 1), 3), np.float64)
    rc = lib().spx_plane_edge_errors(T.ctypes.data, edges.ctypes.data, len(edges), err.ctypes.data)
    if rc != SPX_OK:
        raise SpxError(rc, "spx_plane_edge_errors: bad argument")
    return err[:len(edges)]
