"""CPU, world_size 2 over gloo: frame sharding and the gather of plane lists (the path's only collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sp_slam_b200 import api, sharding


def test_shard_ranges_partition_the_sequence():
    for n, w in ((1000, 8), (1000, 3), (7, 4), (5, 8)):
        r = [sharding.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def _fake_results(rank, n_frames):
    rng = np.random.default_rng(100 + rank)
    hdr = np.zeros(n_frames, api.HEADER_DTYPE)
    hdr["n_planes"] = rng.integers(0, 4 + 3 * rank, n_frames)
    hdr["n_real"] = np.minimum(hdr["n_planes"], 2)
    hdr["first_plane"] = np.concatenate([[0], np.cumsum(hdr["n_planes"])[:-1]])
    n_pl = int(hdr["n_planes"].sum())
    pl = np.zeros(n_pl, api.PLANE_DTYPE)
    pl["coef"] = rng.normal(size=(n_pl, 4)).astype(np.float32)
    pl["n_points"] = rng.integers(500, 30000, n_pl)
    pl["src"] = rank
    return hdr, pl


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hdr, pl = _fake_results(rank, n_frames)
    cap = 64 * n_frames
    buf = np.zeros(cap, api.PLANE_DTYPE)
    buf[: len(pl)] = pl
    out = sharding.gather_records(torch.from_numpy(hdr.view(np.uint8).copy()),
                                  torch.from_numpy(buf.view(np.uint8).copy()),
                                  torch.tensor([len(pl), 0, 0], dtype=torch.int64), to_host=True)
    ok = True
    for k in range(world):
        eh, ep = _fake_results(k, n_frames)
        ok &= np.array_equal(out[k][0], eh) and np.array_equal(out[k][1], ep)
    # the same with an upper bound on the plane count: no read-back of the counts inside the call, validation afterwards
    hint = 10 * n_frames
    hdrs, planes, counts = sharding.gather_records(torch.from_numpy(hdr.view(np.uint8).copy()),
                                                   torch.from_numpy(buf.view(np.uint8).copy()),
                                                   torch.tensor([len(pl), 0, 0], dtype=torch.int64), max_planes_hint=hint)
    ok &= sharding.check_gather(counts, hint) and not sharding.check_gather(counts, 3)
    for k in range(world):
        eh, ep = _fake_results(k, n_frames)
        got = np.frombuffer(planes[k].numpy().tobytes(), dtype=api.PLANE_DTYPE)[: int(counts[k, 0])]
        ok &= np.array_equal(np.frombuffer(hdrs[k].numpy().tobytes(), dtype=api.HEADER_DTYPE), eh) and np.array_equal(got, ep)
    # ragged shards (shard_range on a sequence the world size does not divide): per-rank frame counts travel with the
    # plane counts, headers are padded to frames_cap for the collective
    total = 2 * n_frames - 1
    lo, hi = sharding.shard_range(total, rank, world)
    cap_fr = -(-total // world)
    rh, rp = _fake_results(rank, hi - lo)
    rbuf = np.zeros(cap, api.PLANE_DTYPE)
    rbuf[: len(rp)] = rp
    out = sharding.gather_records(torch.from_numpy(rh.view(np.uint8).copy()), torch.from_numpy(rbuf.view(np.uint8).copy()),
                                  torch.tensor([len(rp), 0, 0], dtype=torch.int64), to_host=True, n_frames=hi - lo, frames_cap=cap_fr)
    for k in range(world):
        a, b = sharding.shard_range(total, k, world)
        eh, ep = _fake_results(k, b - a)
        ok &= len(out[k][0]) == b - a and np.array_equal(out[k][0], eh) and np.array_equal(out[k][1], ep)
    try:     # a rank that owns more frames than the stated bound is an error, not a hang
        sharding.gather_records(torch.from_numpy(rh.view(np.uint8).copy()), torch.from_numpy(rbuf.view(np.uint8).copy()),
                                torch.tensor([len(rp), 0, 0], dtype=torch.int64), to_host=True, n_frames=hi - lo, frames_cap=hi - lo - 1)
        ok = False
    except ValueError:
        pass
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_plane_lists_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
