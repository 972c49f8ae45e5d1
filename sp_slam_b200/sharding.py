"""Frame sharding across the GPUs of one box and the one collective of the path: the gather of plane lists.

The reference is single-process; for offline sequences frames are independent
(PlaneNotSeen only sees the planes of its own frame, /root/reference/src/Frame.cc:1116-1144), so rank r of G owns the
contiguous frame range [r*F/G, (r+1)*F/G) and nothing is exchanged until the per-frame headers (mnRealPlaneNum,
mnPlaneNum) and plane records (mvPlaneCoefficients + cloud sizes) are gathered.  The payload is small (16 B per frame
+ 48 B per plane), so the gather is latency bound; clouds stay on the rank that produced them.
"""
from __future__ import annotations

import numpy as np

from . import api


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous frame range of `rank`; the first n_frames % world ranks get one frame more."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _DevBuf:
    """A raw device range exposed through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _as_tensor(ptr: int, nbytes: int, device):
    import torch
    return torch.as_tensor(_DevBuf(ptr, nbytes), device=device)


def gather_records(hdr, planes_buf, totals, to_host: bool = False, max_planes_hint: int | None = None,
                   n_frames: int | None = None, frames_cap: int | None = None):
    """All-gather of per-rank frame headers and plane records; works on any backend (NCCL on device buffers, gloo on
    host tensors in the CPU tests).

    hdr: uint8 tensor holding this rank's n_frames headers (16 bytes each);  planes_buf: uint8 tensor holding at least
    this rank's plane records (48 bytes each), and at least max-over-ranks records of capacity;  totals: int64[>=3]
    (planes, points, boundary points).  Plane counts differ per rank, so counts are gathered first and the records
    padded to the maximum.

    Ranks may own different numbers of frames (shard_range hands the first n % world ranks one more): n_frames is this
    rank's count (default: what hdr holds) and frames_cap an upper bound valid on EVERY rank (default: n_frames, i.e.
    equal shards; for shard_range use ceil(total / world)).  Headers are padded to frames_cap for the collective; the
    per-rank frame counts travel with the plane counts (4th column of the returned counts).

    max_planes_hint: an upper bound on any rank's plane count known to the caller (e.g. 16 planes per frame).  With it
    the three collectives are enqueued without reading the counts back, so the host never waits for the device inside
    the step; the counts come back as a device tensor and `check_gather` validates them at the caller's next
    synchronisation point (a rank that exceeded the hint => the gather must be repeated without it)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    dev = hdr.device
    hsz = api.HEADER_DTYPE.itemsize
    if n_frames is None:
        n_frames = hdr.numel() // hsz
    if frames_cap is None:
        frames_cap = n_frames
    if n_frames > frames_cap:
        raise ValueError(f"this rank owns {n_frames} frames, more than frames_cap = {frames_cap}")
    if hdr.numel() < n_frames * hsz:
        raise ValueError("header buffer smaller than n_frames")
    meta = torch.cat([totals.reshape(-1)[:3].to(torch.int64), torch.tensor([n_frames], dtype=torch.int64, device=dev)])
    counts = torch.empty(world * 4, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, meta)
    if hdr.numel() >= frames_cap * hsz:
        hdr_pad = hdr[: frames_cap * hsz]
    else:
        hdr_pad = torch.cat([hdr[: n_frames * hsz], torch.zeros(frames_cap * hsz - n_frames * hsz, dtype=torch.uint8, device=dev)])
    hdrs = torch.empty(world * frames_cap * hsz, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(hdrs, hdr_pad.contiguous())
    rec = api.PLANE_DTYPE.itemsize
    if max_planes_hint is not None and not to_host:
        n = max(min(int(max_planes_hint), planes_buf.numel() // rec), 1)
        planes = torch.empty(world * n * rec, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(planes, planes_buf[: n * rec])
        return hdrs.view(world, -1), planes.view(world, -1), counts.view(world, 4)
    counts_h = counts.view(world, 4).cpu()
    if int(counts_h[:, 3].max()) > frames_cap:
        raise ValueError(f"a rank owns {int(counts_h[:, 3].max())} frames, more than frames_cap = {frames_cap}: the ranks disagree on the sharding")
    max_planes = int(counts_h[:, 0].max())
    planes = torch.empty(world * max(max_planes, 1) * rec, dtype=torch.uint8, device=dev)
    if max_planes > 0:
        if planes_buf.numel() < max_planes * rec:
            raise ValueError("plane buffer smaller than the largest rank's plane list")
        dist.all_gather_into_tensor(planes, planes_buf[: max_planes * rec])
    if not to_host:
        return hdrs.view(world, -1), planes.view(world, -1), counts_h
    out = []
    hdrs_h = hdrs.view(world, -1).cpu().numpy()
    planes_h = planes.view(world, -1).cpu().numpy()
    for k in range(world):
        n_pl, n_fr = int(counts_h[k, 0]), int(counts_h[k, 3])
        out.append((np.frombuffer(hdrs_h[k].tobytes(), dtype=api.HEADER_DTYPE)[:n_fr],
                    np.frombuffer(planes_h[k].tobytes(), dtype=api.PLANE_DTYPE)[:n_pl]))
    return out


def check_gather(counts, max_planes_hint: int, frames_cap: int | None = None) -> bool:
    """True when no rank's plane list was longer than the hint the gather was issued with, nor its frame count larger than
    frames_cap (reads the counts: synchronises)."""
    ok = int(counts[:, 0].max()) <= int(max_planes_hint)
    if frames_cap is not None:
        ok = ok and int(counts[:, 3].max()) <= int(frames_cap)
    return ok


def gather_plane_lists(ext: "api.PlaneExtractor", n_frames: int, to_host: bool = False, max_planes_hint: int | None = None,
                       frames_cap: int | None = None):
    """NCCL all-gather of every rank's frame headers and plane records straight from the device buffers of the last
    spx_extract_batch_device (no host round trip for the payload).

    Stream ordering: the extract is asynchronous on the CONTEXT's stream and the collectives are enqueued on torch's
    current stream.  The context is therefore bound to torch's current stream here (spx_set_stream synchronises the old
    stream first, so the extract that produced the buffers is complete), and every later extract of this context runs
    on that stream too -- behind the collectives that read the buffers it will overwrite."""
    import torch

    cur = torch.cuda.current_stream().cuda_stream
    if ext.stream != cur:
        ext.set_stream(cur)
    r = ext.device_results()
    dev = torch.device("cuda", torch.cuda.current_device())
    cap_frames = max(n_frames, frames_cap or 0)
    hdr = _as_tensor(r.frames, cap_frames * api.HEADER_DTYPE.itemsize, dev) if cap_frames <= ext.cfg.max_frames else \
        _as_tensor(r.frames, n_frames * api.HEADER_DTYPE.itemsize, dev)
    totals = _as_tensor(r.totals, 24, dev).view(torch.int64)
    planes_buf = _as_tensor(r.planes, int(r.planes_capacity) * api.PLANE_DTYPE.itemsize, dev)
    return gather_records(hdr, planes_buf, totals, to_host, max_planes_hint, n_frames=n_frames, frames_cap=frames_cap)
