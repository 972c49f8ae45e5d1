#!/usr/bin/env python
"""bench.py -- 640x480 plane-extraction frames/s (BASELINE.json metric) on N B200s, one process per GPU.

A step = one pass of the whole hot path (back-projection .. supposed planes .. packed Frame fields) over one batch of
synthetic depth frames (the box-room orbit of SURVEY.md section 8d, configs[1]); every rank owns its own batch (weak
scaling, frames are independent) and the plane lists are gathered over NCCL once, after the last step (inside the
timed region).  `strong_scaling` (N > 1) is the SAME batch split across the ranks (BASELINE configs[2] as worded).

  value        frames/s, depth batch already resident in HBM, CUDA-event timed, max over ranks
  e2e          frames/s through the C ABI with HOST buffers (spx_extract_batch_compact): pinned host depth -> device,
               kernels, results -> host (the real planes' clouds as ordered inlier index lists)
  e2e_adapter  the same through the C++ host adapter (spx_host::SequencePlanes) until every Frame field of every frame
               is filled (mvPlanePoints / mvBoundaryPoints as 32-byte pcl::PointXYZRGB-layout clouds)
  e2e_full_clouds  the 16-byte-cloud call (spx_extract_batch), round 1's e2e
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
        "k_normals_strip": n * 4 + n * 1 + n * 12 + n * 9,  # the same stage as a strip kernel (TMA-staged depth chunks)
        "k_normals_link_list": 0,                        # frames with NaN / Inf depth only: empty on sensor data
        "k_ccl_merge": n * 1 + n * 4,                    # link bits in, forest touched
        "k_ccl_merge4": n * 1 + n * 4,                   # (the four-pixels-per-thread variant of the same pass)
        "k_ccl_flatten": n * 4 + n * 4 + n * 4,          # forest in/out, counts
        "k_ccl_flatten_runs": n * 4 + n * 4 + n * 4,          # forest in/out, counts
        "k_ccl_frame": n * 1 + n * 4 + n * 4,                 # link bits in, flattened forest out, counts (the forest itself stays in shared memory)
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
    pyoracle.use_native()     # the CPU arm is built for the host it runs on (reference CMakeLists.txt:10-11: -O3 -march=native)
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
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "build": oracle_build_flags(),
                         "sample": f"{n} frames of the workload per step, frame-parallel on {cores} host threads; "
                                   "oracle/ C++ port of the reference's PCL 1.8 path (PCL itself cannot be built here)"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "planes_per_frame": planes / (args.steps * n),
    }
    out.emit(json.dumps(line))


def oracle_build_flags():
    from oracle import pyoracle
    return pyoracle.BUILD_FLAGS


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
    ap.add_argument("--ref-frames", type=int, default=1000, help="frames per step of the CPU arm (default: the whole workload)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank owns --frames frames; strong: --frames frames in total, split across the ranks")
    ap.add_argument("--no-720p", action="store_true", help="skip the short 1280x720 summary (BASELINE configs[3]) of the default line")
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
    warmup_requested = args.warmup
    args.warmup = max(args.warmup, 3)     # timing rule: at least 3 untimed steps; the JSON line states both numbers

    import torch
    import torch.distributed as dist

    from sp_slam_b200 import api
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle
        pyoracle.use_native()      # before anything loads the portable build: the CPU legs run the oracle built for this host

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    rows, cols = (720, 1280) if args.res == "720p" else (ROWS, COLS)
    it = intrinsics(args.res)
    strong = args.scaling == "strong" and world > 1
    if strong:      # the same --frames frames for any N: rank r owns the contiguous range shard_range gives it
        from sp_slam_b200 import sharding
        lo, hi = sharding.shard_range(args.frames, rank, world)
        depth_np = make_frames(args.frames, 0, args.noise, args.res)[lo:hi]
        frames_cap = -(-args.frames // world)
    else:
        depth_np = make_frames(args.frames, rank, args.noise, args.res)
        frames_cap = args.frames
    F = len(depth_np)
    total_frames = args.frames if strong else world * F
    host = torch.from_numpy(np.ascontiguousarray(depth_np)).pin_memory()
    dev = host.cuda(non_blocking=False)
    kw = dict(max_rows=rows, max_cols=cols, device=local_rank, fx=it.fx, fy=it.fy, cx=it.cx, cy=it.cy,
              max_x=float(it.width), max_y=float(it.height))
    ext = api.PlaneExtractor(max_frames=max(F, frames_cap), n_streams=args.streams, **kw)
    # host threads of the library's gathered upload route: half of this rank's share of the host's cores (the library's own default
    # is half of ALL cores, which is right for one process per host only)
    gather_threads = int(os.environ.get("BENCH_GATHER_THREADS", "0")) or max(1, min(16, ((os.cpu_count() or 1) // world) // 2))
    ext.set_gather_threads(gather_threads)
    if not args.profile_groups and (args.streams == 0 or args.streams > 1):
        # the per-kernel table comes from one extra profiled step in which the batch runs as ONE group (kernels back to
        # back on one stream, so their CUDA-event times add up to the step and can be compared with an ncu launch list)
        ext1 = api.PlaneExtractor(max_frames=F, n_streams=1, **kw)
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
    PLANES_HINT = 16   # planes per frame the gather buffers are sized for (validated after the timed region)

    def gather_planes(n_frames):
        """the one collective of the path: plane lists (frame headers + plane records) to every rank over NCCL, once, after
        the last step.  The record gather is padded to a bound of 16 planes per frame so that the three collectives are enqueued
        behind the steps without reading anything back (the counts are validated after the timed region).  Sizing it from the
        gathered counts instead (sharding.gather_plane_lists without the hint) moves 4x fewer bytes but makes the host wait for
        the device before it can launch the record gather: measured 4.73 against 4.32 ms per step on 8 GPUs at 5 steps."""
        if world == 1:
            return
        from sp_slam_b200 import sharding
        gathered[0] = sharding.gather_plane_lists(ext, n_frames, max_planes_hint=PLANES_HINT * frames_cap, frames_cap=frames_cap)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_resident(dev_ptr, n_frames, steps):
        """K passes over the resident batch + ONE final gather of the plane lists, CUDA-event timed on the shared stream."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        n_launch = 0
        for _ in range(steps):
            ext.extract_device(dev_ptr, n_frames, rows, cols)
            n_launch += ext.launches
        gather_planes(n_frames)
        e1.record(stream)
        barrier()
        if world > 1:
            from sp_slam_b200 import sharding
            if not sharding.check_gather(gathered[0][2], PLANES_HINT * frames_cap, frames_cap):
                raise SystemExit("a rank produced more than 16 planes per frame: the gather hint was too small")
        return e0.elapsed_time(e1), n_launch

    for _ in range(args.warmup):
        ext.extract_device(dev.data_ptr(), F, rows, cols)
    gather_planes(F)
    barrier()
    ktimes: dict[str, float] = {}
    kcount: dict[str, int] = {}
    with ClockSampler(local_rank) as clk:
        ms, launches = timed_resident(dev.data_ptr(), F, args.steps)
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
            if name.endswith("_fn"):       # launched through a function pointer that selects the template instance
                name = name[:-3]
            ktimes[name] = ktimes.get(name, 0.0) + t
            kcount[name] = kcount.get(name, 0) + 1
        if ext1 is not None:
            ext1.close()
            ext1 = None

        def timed_host(call, steps):
            """wall clock around `steps` host-buffer calls (copies inside), after 2 untimed ones"""
            r = None
            for _ in range(2):
                r = call()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                r = call()
            barrier()
            return time.perf_counter() - t0, r

        # ---- e2e: host buffers through the C ABI, copies inside the timed region; compact results ----
        # Several ranks on one host share its cores and its memory system: the gathered route then takes from the copy engines what it
        # saves them (2 GPUs: 187 k frames/s with 4 gather threads per rank, 190 k without; 176 k with 8), so it is the single-process setting.
        if world > 1:
            ext.set_upload_mode(2)
        e2e_s, res = timed_host(lambda: ext.extract_batch_compact_ptr(host.data_ptr(), F, rows, cols), args.steps)
        launches_e2e = ext.launches
        xfer = ext.transfer_bytes()   # (uploaded by copies, read in place from the pinned image, copied back) of the last step
        planes_per_frame = float(res.frames["n_planes"].mean())
        overflow = int((res.frames["flags"] & api.SPX_FRAME_OVERFLOW != 0).sum())
        d2h_compact = res.nbytes
        # ---- the same with the sampled rows of every group through the copy engine (no host threads: last round's e2e) ----
        if world == 1:
            ext.set_upload_mode(2)
            e2e_rows_s, _ = timed_host(lambda: ext.extract_batch_compact_ptr(host.data_ptr(), F, rows, cols), args.steps)
            xfer_rows = ext.transfer_bytes()
        else:
            e2e_rows_s, xfer_rows = e2e_s, xfer
        ext.set_upload_mode(0 if world == 1 else 2)
        # ---- the same with 16-byte point clouds back (round 1's e2e) ----
        e2e_full_s, resf = timed_host(lambda: ext.extract_batch_ptr(host.data_ptr(), F, rows, cols), args.steps)
        xfer_full = ext.transfer_bytes()
        # ---- the same with the whole image uploaded (what a pageable caller buffer gets) ----
        ext.set_upload_mode(1)
        e2e_whole_s, _ = timed_host(lambda: ext.extract_batch_compact_ptr(host.data_ptr(), F, rows, cols), args.steps)
        ext.set_upload_mode(0 if world == 1 else 2)
        # ---- through the C++ host adapter until every Frame field of every frame is filled ----
        adapter = None
        try:
            ncpu = os.cpu_count() or 1
            ad = api.SequenceAdapter(api.default_config(max_frames=F, **kw), n_threads=max(1, ncpu // world - 1))
            ad_s, _ = timed_host(lambda: ad.process_ptr(host.data_ptr(), F, rows, cols), args.steps)
            n_pl, n_pt, n_bd, nbytes = ad.summary()
            adc_s, _ = timed_host(lambda: ad.process_clouds_ptr(host.data_ptr(), F, rows, cols), args.steps)
            adapter = {"seconds": ad_s, "seconds_clouds": adc_s, "threads": ad.threads, "planes": n_pl, "points": n_pt, "boundary_points": n_bd,
                       "bytes_filled_per_step": nbytes}
            ad.close()
        except Exception as e:   # noqa: BLE001  (reported, never hidden)
            adapter = {"error": repr(e)}
    # the same end to end from the raw 16-bit depth image (the convertTo of Tracking::GrabImageRGBD fused in)
    e2e16_s = None
    if (rows * cols) % 4 == 0:
        factor = float(np.float32(1.0) / np.float32(5000.0))
        host16 = torch.from_numpy(np.round(np.clip(depth_np, 0, 13.0).astype(np.float64) * 5000.0).astype(np.uint16)).pin_memory()
        e2e16_s, _ = timed_host(lambda: ext.extract_batch_u16_compact_ptr(host16.data_ptr(), F, rows, cols, factor), args.steps)
        xfer16 = ext.transfer_bytes()
    # BASELINE configs[2] as worded: the SAME batch split across the ranks (strong scaling), resident, + the final gather
    strong_ms = None
    if world > 1 and not strong:
        from sp_slam_b200 import sharding
        lo, hi = sharding.shard_range(F, rank, world)
        frames_cap_weak, frames_cap = frames_cap, -(-F // world)
        fb = rows * cols * 4
        for _ in range(2):
            ext.extract_device(dev.data_ptr() + lo * fb, hi - lo, rows, cols)
        strong_ms, _ = timed_resident(dev.data_ptr() + lo * fb, hi - lo, args.steps)
        frames_cap = frames_cap_weak
    # BASELINE configs[4] (tracking loop): one frame at a time through the host-buffer call, as Frame's constructor would
    lat_ms = None
    if rank == 0:
        one = api.PlaneExtractor(max_frames=1, **kw)
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
    ext.close()
    del dev

    # BASELINE configs[3]: a short 1280x720 RealSense-shaped run folded into the default line (rank 0, N = 1)
    cfg720 = None
    if rank == 0 and world == 1 and args.res == "480p" and not args.no_720p:
        cfg720 = measure_720p(stream, not args.no_cpu_baseline)

    t = torch.tensor([ms, e2e_s * 1e3, (e2e16_s or 0.0) * 1e3, e2e_whole_s * 1e3, e2e_full_s * 1e3,
                      (adapter or {}).get("seconds", 0.0) * 1e3, strong_ms or 0.0, (adapter or {}).get("seconds_clouds", 0.0) * 1e3, e2e_rows_s * 1e3],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, e2e16_ms, e2e_whole_ms, e2e_full_ms, ad_ms, strong_ms, adc_ms, e2e_rows_ms = (float(x) for x in t)
    K = args.steps
    value = total_frames * K / (ms * 1e-3)
    per_s = lambda t_ms: total_frames * K / (t_ms * 1e-3) if t_ms else None   # noqa: E731

    if rank == 0:
        peak, peak_src = hbm_peak()
        path_bytes, kbytes, n = algorithmic_bytes_per_frame(rows, cols)
        top = max(ktimes, key=ktimes.get)
        step_kernel_ms = sum(ktimes.values())
        achieved = kbytes.get(top, 0) * F / (ktimes[top] * 1e-3) / 1e9
        traffic, path_traffic = None, None
        tpath = os.path.join(ROOT, "profiles", "r5_traffic_all.json")
        if os.path.exists(tpath) and rows == ROWS and cols == COLS:
            # dram__bytes_read.sum + dram__bytes_write.sum per kernel from one ncu pass over a 1000-frame single-group step
            tj = json.load(open(tpath))
            if top in tj.get("kernels", {}):
                traffic = tj["kernels"][top]["dram_bytes_per_frame"] * F / max(kcount[top], 1)
            path_traffic = tj.get("path_dram_bytes_per_frame")
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
                     "achieved": (value / world) * path_bytes / 1e9, "frac": (value / world) * path_bytes / 1e9 / peak,
                     "measured_dram_bytes_per_frame": path_traffic,
                     "measured_over_algorithmic": (path_traffic / path_bytes) if path_traffic else None},
            "kernels_ms": {k: round(v, 4) for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1])},
            "kernels": [{"kernel": k, "ms": round(v, 4), "algorithmic_bytes_per_frame": kbytes.get(k, 0),
                         "achieved_gbs": round(kbytes.get(k, 0) * F / (v * 1e-3) / 1e9, 1) if v > 0 else None,
                         "frac": round(kbytes.get(k, 0) * F / (v * 1e-3) / 1e9 / peak, 4) if v > 0 else None}
                        for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1])],
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_leg(depth_np, min(F, args.cpu_sample), args.res)
        line = {
            "metric": METRIC if args.res == "480p" else "1280x720 plane-extraction frames/s",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "warmup_requested": warmup_requested,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "frames_per_step_per_gpu": F, "frames_per_step_total": total_frames,
                       "noise": args.noise,
                       "cloud_dis": 3, "organized_cloud": n, "l2": "inputs larger than L2 (depth batch "
                       f"{F * rows * cols * 4 / 1e6:.0f} MB per GPU)", "parallelism": f"frame-sharded x{world}",
                       "planes_per_frame": planes_per_frame, "overflow_frames": overflow,
                       "collective": "none" if world == 1 else "one NCCL all-gather of the plane lists after the last step, inside the timed region"},
            "e2e": {"value": per_s(e2e_ms), "unit": UNIT, "h2d_bytes_per_step": xfer[0] + xfer[1],
                    "d2h_bytes_per_step": xfer[2], "ms_per_step": e2e_ms / K,
                    "h2d_copied": xfer[0], "h2d_read_in_place": xfer[1], "result_bytes": d2h_compact, "gpu_launches_per_step": launches_e2e,
                    "host_gather_threads": gather_threads if world == 1 else 0,
                    "upload": "two routes (copy engine: sampled rows of the first groups; host threads: gathered samples of the last groups)" if world == 1
                              else "sampled rows through the copy engine only (several ranks share the host's cores and memory system)",
                    "note": "spx_extract_batch_compact on the pinned CV_32F batch, two upload routes at once: the first frame groups' sampled rows "
                            "(every Cloud.Dis-th) go through the copy engine, one strided copy per group, while host threads inside the library "
                            "gather the organized cloud's samples (every Cloud.Dis-th row AND column) of the last groups into a pinned staging "
                            "buffer, of which only those samples are uploaded; the sectors of the 21x21 full-resolution windows the "
                            "border tests read are fetched from the pinned image over PCIe by k_border_fetch (h2d_read_in_place); back come "
                            "frame headers, plane records, boundary clouds, the supposed planes' clouds and -- instead of the real planes' "
                            "clouds -- their ordered inlier index lists, from which the host adapter rebuilds pcl::PointXYZRGB bit-exactly"},
            "e2e_adapter": None if not adapter else (adapter if "error" in adapter else {
                "value": per_s(ad_ms), "unit": UNIT, "ms_per_step": ad_ms / K, "host_threads": adapter["threads"],
                "bytes_filled_per_step": adapter["bytes_filled_per_step"], "points_per_step": adapter["points"] + adapter["boundary_points"],
                "note": "spx_host::SequencePlanes (C++): the same call, then every Frame field of every frame filled -- mvPlanePoints / "
                        "mvBoundaryPoints as 32-byte pcl::PointXYZRGB-layout clouds in pooled storage, mvPlaneCoefficients -- by host threads "
                        "that start on a frame group as soon as it is on the host (spx_set_group_callback)",
                "clouds_transfer": {"value": per_s(adc_ms), "unit": UNIT, "ms_per_step": adc_ms / K,
                                    "note": "the same adapter fed by spx_extract_batch: the real planes' clouds cross PCIe as 16-byte points and "
                                            "are only widened on the host (more bytes on the link, less host arithmetic)"}}),
            "e2e_sampled_rows": {"value": per_s(e2e_rows_ms), "unit": UNIT, "h2d_bytes_per_step": xfer_rows[0] + xfer_rows[1],
                                 "d2h_bytes_per_step": xfer_rows[2], "ms_per_step": e2e_rows_ms / K,
                                 "note": "the compact call with every group's sampled rows through the copy engine and no host threads "
                                         "(spx_set_upload_mode 2; the e2e of the previous bench lines)"},
            "e2e_full_clouds": {"value": per_s(e2e_full_ms), "unit": UNIT, "h2d_bytes_per_step": xfer_full[0] + xfer_full[1],
                                "d2h_bytes_per_step": xfer_full[2], "ms_per_step": e2e_full_ms / K,
                                "note": "spx_extract_batch: all clouds back as 16-byte points (round 1's e2e)"},
            "e2e_whole_image": {"value": per_s(e2e_whole_ms), "unit": UNIT,
                                "h2d_bytes_per_step": F * rows * cols * 4, "ms_per_step": e2e_whole_ms / K,
                                "note": "the compact call with the whole image uploaded (spx_set_upload_mode 1; what a pageable buffer gets)"},
            "e2e_u16": None if not e2e16_ms else {
                "value": per_s(e2e16_ms), "unit": UNIT, "h2d_bytes_per_step": xfer16[0] + xfer16[1], "d2h_bytes_per_step": xfer16[2],
                "ms_per_step": e2e16_ms / K,
                "note": "host input = the raw CV_16U depth image, DepthMapFactor conversion on the device (spx_extract_batch_u16_compact)"},
            "strong_scaling": None if not strong_ms else {
                "value": F * K / (strong_ms * 1e-3), "unit": UNIT, "frames_total": F, "ms_per_step": strong_ms / K,
                "note": f"the same {F}-frame batch split across the {world} ranks (shard_range), resident, K steps + one final "
                        "gather of the plane lists; BASELINE configs[2] as worded"},
            "gpu_launches": launches,
            "latency_ms_per_frame_in_batch": ms / args.steps / F,
            "single_frame_latency_ms": lat_ms,
            "roofline": roofline,
            "config4_720p": cfg720,
            "next_rows": next_rows,
            "cpu_baseline": cpu,
            "clocks": clk.summary(),
        }
        out.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline_leg(depth_np, ns, res):
    """The oracle port on all host threads, bounded sample: ~10-30 s of CPU work over all threads."""
    cores = os.cpu_count() or 1
    oracle_fps(depth_np[: max(cores, 8)], cores, res)     # thread start-up, page faults
    reps, dt, tp, ts = 0, 0.0, 0.0, 0.0
    while reps < 3 or (dt < 2.0 and reps < 12):
        _, d1, _, a, b = oracle_fps(depth_np[:ns], cores, res)
        reps += 1; dt += d1; tp += a; ts += b
    fps = reps * ns / dt
    return {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "build": oracle_build_flags(),
            "sample": f"first {ns} frames of the workload x {reps} passes, frame-parallel on {cores} host threads, "
                      f"{dt:.1f} s wall / {tp + ts:.1f} s of CPU work; "
                      f"oracle/ C++ port of the reference's PCL 1.8 path; per-frame 1-core time "
                      f"{1e3 * (tp + ts) / (ns * reps):.2f} ms (plane {1e3 * tp / (ns * reps):.2f} + supposed {1e3 * ts / (ns * reps):.2f})"}


def measure_720p(stream, with_cpu, n_frames=120, steps=3):
    """BASELINE configs[3]: 1280x720 noisy RealSense-shaped clutter frames (many small planes, edge-generated supposed planes):
    resident frames/s, end to end (compact), HBM-roofline fraction of the path, overflow frames, CPU port beside it."""
    import torch
    from sp_slam_b200 import api, scenes
    rows, cols = 720, 1280
    it = scenes.REALSENSE
    d = scenes.realsense_sequence(n_frames)
    d = np.stack([scenes.add_noise(d[k], k, "realsense") for k in range(n_frames)])
    host = torch.from_numpy(d).pin_memory()
    dev = host.cuda()
    ext = api.PlaneExtractor(max_frames=n_frames, max_rows=rows, max_cols=cols, fx=it.fx, fy=it.fy, cx=it.cx, cy=it.cy,
                             max_x=float(it.width), max_y=float(it.height), device=torch.cuda.current_device())
    ext.set_stream(stream.cuda_stream)
    for _ in range(3):
        ext.extract_device(dev.data_ptr(), n_frames, rows, cols)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        ext.extract_device(dev.data_ptr(), n_frames, rows, cols)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    for _ in range(2):
        res = ext.extract_batch_compact_ptr(host.data_ptr(), n_frames, rows, cols)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = ext.extract_batch_compact_ptr(host.data_ptr(), n_frames, rows, cols)
    e2e_s = (time.perf_counter() - t0) / steps
    xfer = ext.transfer_bytes()
    peak, _ = hbm_peak()
    path_bytes, _, n = algorithmic_bytes_per_frame(rows, cols)
    value = n_frames / (ms * 1e-3)
    out = {"workload": f"{n_frames}-frame synthetic 1280x720 RealSense-shaped clutter sequence with sensor noise, Plane.MinSize 500",
           "value": value, "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "e2e": {"value": n_frames / e2e_s, "unit": UNIT, "h2d_bytes_per_step": xfer[0] + xfer[1], "d2h_bytes_per_step": xfer[2]},
           "organized_cloud": n, "planes_per_frame": float(res.frames["n_planes"].mean()),
           "real_planes_per_frame": float(res.frames["n_real"].mean()),
           "overflow_frames": int((res.frames["flags"] & api.SPX_FRAME_OVERFLOW != 0).sum()),
           "roofline_path": {"algorithmic_bytes_per_frame": path_bytes, "achieved": value * path_bytes / 1e9,
                             "frac": value * path_bytes / 1e9 / peak}}
    ext.close()
    if with_cpu:
        cores = os.cpu_count() or 1
        oracle_fps(d[:cores], cores, "720p")
        fps, dt, _, tp, ts = oracle_fps(d, cores, "720p")
        out["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "build": oracle_build_flags(),
                               "sample": f"the same {n_frames} frames, one pass, frame-parallel on {cores} host threads, {dt:.1f} s wall"}
    return out


if __name__ == "__main__":
    main()
