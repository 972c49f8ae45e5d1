#!/usr/bin/env python
"""bench.py -- 640x480 plane-extraction frames/s (BASELINE.json metric) on N B200s, one process per GPU.

A step = one pass of the whole hot path (back-projection .. supposed planes .. packed Frame fields) over one batch of
synthetic depth frames (the box-room orbit of SURVEY.md section 8d, configs[1]); every rank owns its own batch (weak
scaling, frames are independent) and the plane lists are gathered over NCCL at the end of every step.

  value        frames/s, depth batch already resident in HBM, CUDA-event timed, max over ranks
  e2e          frames/s through the C ABI with HOST buffers: pinned host depth -> device, kernels, Frame fields -> host
  roofline     the kernel with the largest share of the step, its algorithmic bytes (table below) over its
               CUDA-event duration, against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline the CPU oracle (a port of the reference's PCL path; oracle/) on the host cores, bounded sample

`--impl reference` times that CPU oracle on all host threads instead (the reference's own PCL build cannot be compiled
here: PCL / Eigen / Boost / OpenCV C++ are absent).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# The library runs a batch as frame groups on their own CUDA streams (+ one upload and one download stream); streams beyond
# the driver's hardware-queue count share a queue and serialise.  Read at CUDA context creation, so set before torch starts.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "640x480 plane-extraction frames/s"
UNIT = "frames/s"
ROWS, COLS = 480, 640


def algorithmic_bytes_per_frame(rows=ROWS, cols=COLS, dis=3):
    """Compulsory bytes per frame: whole path and per kernel (inputs read once + outputs written once,
    upper bound = every pixel an inlier).  N = organized-cloud size."""
    w, h = -(-cols // dis), -(-rows // dis)
    n = w * h
    path = rows * cols * 4 + n * 4 + n * 4 + n * 16 + 65536
    k = {
        "k_edge_chamfer": n * 4 + n * 1,                 # one depth sample per organized pixel in, window size out
        "k_normals_link": n * 4 + n * 1 + n * 12 + n * 9, # depth sample + window size in; x y z, link bits, forest, counts out
        "k_ccl_merge": n * 1 + n * 4,                    # link bits in, forest touched
        "k_ccl_merge4": n * 1 + n * 4,                   # (the four-pixels-per-thread variant of the same pass)
        "k_ccl_flatten": n * 4 + n * 4 + n * 4,          # forest in/out, counts
        "k_ccl_flatten_runs": n * 4 + n * 4 + n * 4,          # forest in/out, counts
        "k_ccl_rank": n * 8 + n * 2 + n * 8,             # forest + sizes in, index lists + positions out (upper bound)
        "k_ccl_label": n * 8,
        "k_moments_fit": n * 4 + n * 12,                 # index list + xyz of the members in
        "k_models": 8192,
        "k_pid_init": n * 4 + n * 1,
        "k_pid_init4": n * 4 + n * 1,
        "k_refine": 2 * 2 * n + n * 12 + n * 4,          # plane ids in/out twice, xyz of free pixels, positions
        "k_refine2": 2 * 2 * n + n * 12 + n * 4,
        "k_contour": n * 1 + 16384,                      # plane-id map in, contour indices out
        "k_postfilter": 8192,
        "k_lines": 4 * 4096 * 16,                        # <= 4 rounds over a contour of a few thousand points
        "k_border": 2 * 160000,                          # 20x20 full-resolution depth windows of the line points
        "k_supposed": 8192,
        "k_scan_frames": 64,
        "k_emit_records": 8192,
        "k_pack_points": n * 17 + n * 16,                # pid pos xyz in, 16-byte points out
        "k_pack_contours": 4096 * (4 + 12 + 16),
        "k_pack_supposed": 2 * 2500 * 16,
        "k_convert_u16": rows * cols * 6,
    }
    return path, k, n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_frames(n_frames: int, rank: int, noise: str, res: str):
    from sp_slam_b200 import scenes
    start = (rank * 125) % 1000
    if res == "720p":
        d = scenes.realsense_sequence(n_frames, start=start)
        model = "realsense"
    else:
        d = scenes.boxroom_sequence(n_frames, start=start)
        model = "kinect"
    if noise != "none":
        for k in range(n_frames):
            d[k] = scenes.add_noise(d[k], start + k, model)
    return d


def oracle_fps(depth: np.ndarray, n_threads: int, res: str):
    from oracle import pyoracle
    cfg = oracle_config(res)
    t0 = time.perf_counter()
    nr, na, tp, ts = pyoracle.run_batch(depth, n_threads, cfg)
    dt = time.perf_counter() - t0
    return len(depth) / dt, dt, int(na.sum()), tp, ts


def measure_next_rows(ext, dev, host, F, rows, cols, stream, with_cpu):
    """N4: device-side voxel downsampling (leaf 0.01 / 0.05) of every contour of the batch, between the extract and the
    fetch.  N1: per-frame Map::AssociatePlanesByBoundary against a map made of the planes of the first 40 frames."""
    import torch
    from sp_slam_b200 import api
    out = {}
    ext.set_profile(False)
    ext.extract_device(dev.data_ptr(), F, rows, cols)
    base = ext.fetch()
    n_bnd = len(base.boundary)
    for leaf in (0.01, 0.05):
        ts = []
        for _ in range(3):
            ext.extract_device(dev.data_ptr(), F, rows, cols)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ext.voxel_downsample_results(leaf, 1)
            ts.append(time.perf_counter() - t0)
        down = ext.fetch()
        rec = {"leaf_m": leaf, "clouds": int(len(base.planes)), "points_in": n_bnd, "points_out": int(len(down.boundary)),
               "ms": 1e3 * min(ts), "mpoints_per_s": n_bnd / min(ts) / 1e6}
        if with_cpu:
            from oracle import pyoracle
            t0 = time.perf_counter()
            k = 0
            npts = 0
            while time.perf_counter() - t0 < 1.0 and k < len(base.planes):
                q = base.planes[k]
                pyoracle.voxel_grid(base.boundary[q["boundary_off"]:q["boundary_off"] + q["n_boundary"]], leaf)
                npts += int(q["n_boundary"]); k += 1
            dt = time.perf_counter() - t0
            rec["cpu_port_mpoints_per_s_1core"] = npts / dt / 1e6
        out[f"voxel_grid_contours_leaf{leaf}"] = rec
    fps = [base.frame(k) for k in range(min(F, 140))]
    map_fps = fps[:40]
    map_w = np.concatenate([fp.mvPlaneCoefficients for fp in map_fps])
    bnds = [b for fp in map_fps for b in fp.mvBoundaryPoints]
    pm = api.PlaneMap(ext)
    pm.upload(map_w, bnds)
    ts, tc, n_assoc = [], [], 0
    for fp in fps[40:]:
        t0 = time.perf_counter()
        a, v, p, d = pm.associate(fp.mvPlaneCoefficients)
        ts.append(time.perf_counter() - t0)
        n_assoc += int((a >= 0).sum())
        if with_cpu:
            from oracle import pyoracle
            t0 = time.perf_counter()
            pyoracle.associate_planes(fp.mvPlaneCoefficients, map_w, bnds)
            tc.append(time.perf_counter() - t0)
    # MapPlane::UpdateBoundary from the device results of the last extract (frame 60's planes onto map planes 0..)
    ext.extract_device(dev.data_ptr(), F, rows, cols)
    torch.cuda.synchronize()
    tu = []
    k60 = int(np.argmax([fp.mnPlaneNum for fp in fps]))
    fp60 = fps[k60]
    T = np.eye(4); T[:3, 3] = (0.1, -0.2, 0.05)
    for rep_ in range(20):
        for pl in range(fp60.mnPlaneNum):
            t0 = time.perf_counter()
            pm.update_boundary_from_result(pl % len(map_w), T, k60, pl, len(fp60.mvBoundaryPoints[pl]))
            tu.append(time.perf_counter() - t0)
    pm.close()
    out["map_boundary_update"] = {"calls": len(tu), "ms_per_call_median": 1e3 * float(np.median(tu)) if tu else None,
                                  "note": "MapPlane::UpdateBoundary: transform of a frame plane's contour into the map, source read in "
                                          "place in the device result arena (spx_map_update_boundary_from_result)"}
    out["plane_association"] = {"map_planes": int(len(map_w)), "map_boundary_points": int(sum(len(b) for b in bnds)),
                                "frames": len(ts), "associated_planes": n_assoc,
                                "ms_per_frame_median": 1e3 * float(np.median(ts)),
                                "cpu_port_ms_per_frame_median": (1e3 * float(np.median(tc))) if tc else None,
                                "note": "host planes in, indices out, one call per frame (Map::AssociatePlanesByBoundary); "
                                        "the CPU figure includes the ctypes call of the oracle"}
    return out


def intrinsics(res):
    from sp_slam_b200 import scenes
    return scenes.REALSENSE if res == "720p" else scenes.TUM1


def oracle_config(res):
    from oracle import pyoracle
    it = intrinsics(res)
    return pyoracle.default_config(fx=it.fx, fy=it.fy, cx=it.cx, cy=it.cy, max_x=float(it.width), max_y=float(it.height))


def run_reference(args, rank, world, out):
    """CPU arm: the oracle port of the reference's PCL path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = min(args.frames, args.ref_frames)
    depth = make_frames(n, 0, args.noise, args.res)
    for _ in range(args.warmup):
        oracle_fps(depth[: max(cores, 8)], cores, args.res)
    t0 = time.perf_counter()
    planes = 0
    for _ in range(args.steps):
        _, _, p, _, _ = oracle_fps(depth, cores, args.res)
        planes += p
    dt = time.perf_counter() - t0
    fps = args.steps * n / dt
    line = {
        "impl": "reference", "metric": METRIC if args.res == "480p" else "1280x720 plane-extraction frames/s",
        "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "frames_per_step_per_gpu": args.frames, "noise": args.noise,
                   "cloud_dis": 3, "sample_frames_per_step": n},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} frames of the workload per step, frame-parallel on {cores} host threads; "
                                   "oracle/ C++ port of the reference's PCL 1.8 path (PCL itself cannot be built here)"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "planes_per_frame": planes / (args.steps * n),
    }
    out.emit(json.dumps(line))


def workload_name(args):
    if args.res == "720p":
        return f"{args.frames}-frame synthetic 1280x720 RealSense-shaped clutter sequence, batched plane extraction"
    return f"{args.frames}-frame synthetic 640x480 box-room orbit (BASELINE configs[1]), batched plane extraction"


class _StdoutToStderr:
    """Everything libraries print on fd 1 (NCCL's version banner, ...) goes to stderr; the JSON line is the only thing
    that reaches the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        sys.stdout.flush()
        os.write(self._saved, (line + "\n").encode())

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def main():
    with _StdoutToStderr() as out:
        _main(out)


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1000, help="frames per step and per GPU")
    ap.add_argument("--ref-frames", type=int, default=500, help="frames per step of the CPU arm")
    ap.add_argument("--cpu-sample", type=int, default=1000, help="frames of the cpu_baseline leg")
    ap.add_argument("--noise", default="none", choices=["none", "sensor"])
    ap.add_argument("--res", default="480p", choices=["480p", "720p"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-groups", action="store_true", help="take the kernel table from the multi-group timed steps")
    ap.add_argument("--streams", type=int, default=0, help="internal streams (frame groups) per context; 0 = library default")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    from sp_slam_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    rows, cols = (720, 1280) if args.res == "720p" else (ROWS, COLS)
    F = args.frames
    it = intrinsics(args.res)
    depth_np = make_frames(F, rank, args.noise, args.res)
    host = torch.from_numpy(depth_np).pin_memory()
    dev = host.cuda(non_blocking=False)
    ext = api.PlaneExtractor(max_frames=F, max_rows=rows, max_cols=cols, device=local_rank, fx=it.fx, fy=it.fy, cx=it.cx,
                             cy=it.cy, max_x=float(it.width), max_y=float(it.height), n_streams=args.streams)
    if not args.profile_groups and (args.streams == 0 or args.streams > 1):
        # the per-kernel table comes from one extra profiled step in which the batch runs as ONE group (kernels back to
        # back on one stream, so their CUDA-event times add up to the step and can be compared with an ncu launch list)
        ext1 = api.PlaneExtractor(max_frames=F, max_rows=rows, max_cols=cols, device=local_rank, fx=it.fx, fy=it.fy,
                                  cx=it.cx, cy=it.cy, max_x=float(it.width), max_y=float(it.height), n_streams=1)
    else:
        ext1 = None
    # a real (non-default) stream shared by torch and the library, so torch.cuda.Event brackets the kernels
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    ext.set_stream(stream.cuda_stream)
    ext.set_profile(ext1 is None)
    if ext1 is not None:
        ext1.set_stream(stream.cuda_stream)
        ext1.set_profile(True)

    gathered = [None]

    def gather_planes():
        """the one collective of the path: plane lists (frame headers + plane records) to every rank over NCCL"""
        if world == 1:
            return
        from sp_slam_b200 import sharding
        # an upper bound of 16 planes per frame lets the three collectives be enqueued without reading the counts back
        # (no host wait inside the step); the counts of the last step are validated after the timed region
        gathered[0] = sharding.gather_plane_lists(ext, F, max_planes_hint=16 * F)

    def step_device():
        ext.extract_device(dev.data_ptr(), F, rows, cols)
        gather_planes()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    ktimes: dict[str, float] = {}
    kcount: dict[str, int] = {}
    launches = 0
    with ClockSampler(local_rank) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
            launches += ext.launches
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            from sp_slam_b200 import sharding
            if not sharding.check_gather(gathered[0][2], 16 * F):
                raise SystemExit("a rank produced more than 16 planes per frame: the gather hint was too small")
        ms_one = None
        if ext1 is not None:
            for _ in range(2):
                ext1.extract_device(dev.data_ptr(), F, rows, cols)
            barrier()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(stream)
            ext1.extract_device(dev.data_ptr(), F, rows, cols)
            p1.record(stream)
            barrier()
            ms_one = p0.elapsed_time(p1)
        for name, t in (ext1 or ext).kernel_times():   # CUDA events around every launch of the profiled step
            name = name.split("<")[0]
            ktimes[name] = ktimes.get(name, 0.0) + t
            kcount[name] = kcount.get(name, 0) + 1
        # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
        res = None
        for _ in range(2):
            res = ext.extract_batch_ptr(host.data_ptr(), F, rows, cols)
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(args.steps):
            res = ext.extract_batch_ptr(host.data_ptr(), F, rows, cols)
            d2h += res.frames.nbytes + res.planes.nbytes + res.points.nbytes + res.boundary.nbytes + 24
            launches_e2e = ext.launches
        barrier()
        e2e_s = time.perf_counter() - t0
        xfer = ext.transfer_bytes()   # (uploaded by copies, read in place from the pinned image, copied back) of the last step
        # the same with the whole image uploaded (what a pageable caller buffer gets)
        ext.set_upload_mode(1)
        for _ in range(2):
            ext.extract_batch_ptr(host.data_ptr(), F, rows, cols)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ext.extract_batch_ptr(host.data_ptr(), F, rows, cols)
        barrier()
        e2e_whole_s = time.perf_counter() - t0
        ext.set_upload_mode(0)
    # the same end to end from the raw 16-bit depth image (the convertTo of Tracking::GrabImageRGBD fused in)
    e2e16_s = None
    if (rows * cols) % 4 == 0:
        factor = float(np.float32(1.0) / np.float32(5000.0))
        host16 = torch.from_numpy(np.round(np.clip(depth_np, 0, 13.0).astype(np.float64) * 5000.0).astype(np.uint16)).pin_memory()
        for _ in range(2):
            ext.extract_batch_u16_ptr(host16.data_ptr(), F, rows, cols, factor)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ext.extract_batch_u16_ptr(host16.data_ptr(), F, rows, cols, factor)
        barrier()
        e2e16_s = time.perf_counter() - t0
        xfer16 = ext.transfer_bytes()
    # BASELINE configs[4] (tracking loop): one frame at a time through the host-buffer call, as Frame's constructor would
    lat_ms = None
    if rank == 0:
        one = api.PlaneExtractor(max_frames=1, max_rows=rows, max_cols=cols, device=local_rank, fx=it.fx, fy=it.fy, cx=it.cx,
                                 cy=it.cy, max_x=float(it.width), max_y=float(it.height))
        def lat(mode):
            one.set_upload_mode(mode)
            ts = []
            for k in range(min(F, 60)):
                t1 = time.perf_counter()
                one.extract_batch_ptr(host[k].data_ptr(), 1, rows, cols)
                ts.append((time.perf_counter() - t1) * 1e3)
            ts = np.array(ts[10:])
            return {"median": float(np.median(ts)), "p95": float(np.percentile(ts, 95)), "max": float(ts.max()), "frames": len(ts)}
        lat_ms = lat(0)
        lat_ms["budget_ms"] = 33.3
        lat_ms["whole_image_upload"] = lat(1)
        lat_ms["sparse_upload"] = lat(2)
        # the loop of BASELINE configs[4] for the plane landmarks: extraction of one frame, association of its planes against
        # a device-resident map (Map::AssociatePlanesByBoundary), pose-only optimisation with the plane edges
        # (Optimizer::PoseOptimization's plane part, host code as in the reference), boundary update of the associated map
        # planes (MapPlane::UpdateBoundary).  The ORB point pipeline of the same loop is outside this repository.
        one.set_upload_mode(0)
        first = one.extract_batch_ptr(host[0].data_ptr(), 1, rows, cols, copy=True).frame(0)
        pm = api.PlaneMap(one)
        pm.upload(first.mvPlaneCoefficients, first.mvBoundaryPoints)
        eye = np.eye(4)
        tl, n_up, n_edges = [], 0, 0
        for k in range(1, min(F, 61)):
            t1 = time.perf_counter()
            fr = one.extract_batch_ptr(host[k].data_ptr(), 1, rows, cols).frame(0)
            a, v, p, dd = pm.associate(fr.mvPlaneCoefficients)
            kinds, map_pl, frame_pl = [], [], []
            for kind, idx in ((0, a), (1, p), (2, v)):
                for i, j in enumerate(idx):
                    if j >= 0:
                        kinds.append(kind); map_pl.append(first.mvPlaneCoefficients[j]); frame_pl.append(fr.mvPlaneCoefficients[i])
            if kinds:
                api.pose_optimize_planes(eye, api.plane_edges(kinds, map_pl, frame_pl))
                n_edges += len(kinds)
            for i, j in enumerate(a):
                if j >= 0:
                    pm.update_boundary_from_result(int(j), eye, 0, i, len(fr.mvBoundaryPoints[i]))
                    n_up += 1
            tl.append((time.perf_counter() - t1) * 1e3)
        pm.close()
        tl = np.array(tl[10:])
        lat_ms["tracking_loop"] = {"median": float(np.median(tl)), "p95": float(np.percentile(tl, 95)), "frames": len(tl),
                                   "boundary_updates": n_up, "plane_edges": n_edges,
                                   "note": "extract + associate + pose optimisation with the plane edges (host) + boundary updates "
                                           "per frame, through the Python binding"}
        one.close()
    # SURVEY 8(f) rows either side of the path (N4 VoxelGrid of the contours, N1 plane association), rank 0 at N = 1
    next_rows = None
    if rank == 0 and world == 1:
        next_rows = measure_next_rows(ext, dev, host, F, rows, cols, stream, not args.no_cpu_baseline)
    planes_per_frame = float(res.frames["n_planes"].mean())
    overflow = int((res.frames["flags"] != 0).sum())

    t = torch.tensor([ms, e2e_s * 1e3, (e2e16_s or 0.0) * 1e3, e2e_whole_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, e2e16_ms, e2e_whole_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    value = world * F * args.steps / (ms * 1e-3)
    e2e_val = world * F * args.steps / (e2e_ms * 1e-3)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        path_bytes, kbytes, n = algorithmic_bytes_per_frame(rows, cols)
        top = max(ktimes, key=ktimes.get)
        step_kernel_ms = sum(ktimes.values())
        achieved = kbytes.get(top, 0) * F / (ktimes[top] * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath):   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, scaled to this launch
            tj = json.load(open(tpath)).get(top)
            if tj and rows == ROWS and cols == COLS:
                traffic = tj["dram_bytes_per_frame"] * F / max(kcount[top], 1)
        roofline = {
            "bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src,
            "kernel_ms_per_step": ktimes[top], "kernel_launches_per_step": kcount[top],
            "kernel_ms_per_launch": ktimes[top] / kcount[top], "kernel_share_of_step": ktimes[top] / step_kernel_ms,
            "profiled_step_ms": ms_one,
            "note": "kernel durations: CUDA events around every launch of one extra step run as a single frame group (kernels "
                    "back to back on one stream); `value` is timed with the batch cut into frame groups on internal streams",
            "algorithmic_bytes_per_frame": kbytes.get(top, 0),
            "path": {"algorithmic_bytes_per_frame": path_bytes,
                     "achieved": (value / world) * path_bytes / 1e9, "frac": (value / world) * path_bytes / 1e9 / peak},
            "kernels_ms": {k: round(v, 4) for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1])},
            "kernels": [{"kernel": k, "ms": round(v, 4), "algorithmic_bytes_per_frame": kbytes.get(k, 0),
                         "achieved_gbs": round(kbytes.get(k, 0) * F / (v * 1e-3) / 1e9, 1) if v > 0 else None,
                         "frac": round(kbytes.get(k, 0) * F / (v * 1e-3) / 1e9 / peak, 4) if v > 0 else None}
                        for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1])],
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            ns = min(F, args.cpu_sample)
            oracle_fps(depth_np[: max(cores, 8)], cores, args.res)     # thread start-up, page faults
            reps, dt, tp, ts = 0, 0.0, 0.0, 0.0
            while reps < 3 or (dt < 2.0 and reps < 12):                # ~10-30 s of CPU work over all threads
                _, d1, _, a, b = oracle_fps(depth_np[:ns], cores, args.res)
                reps += 1; dt += d1; tp += a; ts += b
            fps = reps * ns / dt
            cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {ns} frames of the workload x {reps} passes, frame-parallel on {cores} host threads, "
                             f"{dt:.1f} s wall / {tp + ts:.1f} s of CPU work; "
                             f"oracle/ C++ port of the reference's PCL 1.8 path; per-frame 1-core time "
                             f"{1e3 * (tp + ts) / (ns * reps):.2f} ms (plane {1e3 * tp / (ns * reps):.2f} + supposed {1e3 * ts / (ns * reps):.2f})"}
        line = {
            "metric": METRIC if args.res == "480p" else "1280x720 plane-extraction frames/s",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "frames_per_step_per_gpu": F, "noise": args.noise,
                       "cloud_dis": 3, "organized_cloud": n, "l2": "inputs larger than L2 (depth batch "
                       f"{F * rows * cols * 4 / 1e6:.0f} MB per GPU)", "parallelism": f"frame-sharded x{world}",
                       "planes_per_frame": planes_per_frame, "overflow_frames": overflow},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": xfer[0] + xfer[1],
                    "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms / args.steps,
                    "h2d_copied": xfer[0], "h2d_read_in_place": xfer[1],
                    "note": "spx_extract_batch on the pinned CV_32F batch: the rows the organized cloud samples (every Cloud.Dis-th) "
                            "are uploaded with one strided copy per frame group, the sectors of the 21x21 full-resolution windows the "
                            "border tests read are fetched from the pinned image over PCIe by k_border_fetch (h2d_read_in_place); "
                            "all Frame fields (planes + clouds) come back"},
            "e2e_whole_image": {"value": world * F * args.steps / (e2e_whole_ms * 1e-3), "unit": UNIT,
                                "h2d_bytes_per_step": F * rows * cols * 4, "ms_per_step": e2e_whole_ms / args.steps,
                                "note": "the same call with the whole image uploaded (spx_set_upload_mode 1; what a pageable buffer gets)"},
            "e2e_u16": None if not e2e16_ms else {
                "value": world * F * args.steps / (e2e16_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": xfer16[0] + xfer16[1],
                "ms_per_step": e2e16_ms / args.steps,
                "note": "host input = the raw CV_16U depth image, DepthMapFactor conversion on the device (spx_extract_batch_u16)"},
            "gpu_launches": launches,
            "latency_ms_per_frame_in_batch": ms / args.steps / F,
            "single_frame_latency_ms": lat_ms,
            "roofline": roofline,
            "next_rows": next_rows,
            "cpu_baseline": cpu,
            "clocks": clk.summary(),
        }
        out.emit(json.dumps(line))
    ext.close()
    if ext1 is not None:
        ext1.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
