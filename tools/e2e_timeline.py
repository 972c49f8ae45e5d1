"""Per-group timeline of one host-input step (development aid): python tools/e2e_timeline.py [frames] [streams] [mode]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sp_slam_b200 import api, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
st = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sup = int(sys.argv[4]) if len(sys.argv) > 4 else 1
d = scenes.boxroom_sequence(n)
U16 = int(os.environ.get('SPX_TL_U16', '0'))
if U16:
    d = np.round(np.clip(d, 0, 13.0).astype(np.float64) * 5000.0).astype(np.uint16)
host = torch.from_numpy(d).pin_memory()
dev = torch.empty_like(host, device="cuda")
if not os.environ.get('SPX_TL_NORATES') and not U16:
    # raw copy rates
    def timeit(fn, reps=3):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3
    print("H2D whole 1.23GB ms", timeit(lambda: dev.copy_(host, non_blocking=True)))
    hs = host.view(n * 160, 3, 640)[:, 0, :]
    ds = torch.empty(n * 160, 640, device="cuda")
    print("H2D strided rows 0.41GB ms", timeit(lambda: ds.copy_(hs, non_blocking=True)))
    back = torch.empty(437_000_000 // 4, dtype=torch.float32).pin_memory()
    src = torch.empty(437_000_000 // 4, dtype=torch.float32, device="cuda")
    print("D2H 437MB ms", timeit(lambda: back.copy_(src, non_blocking=True)))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def both():
        with torch.cuda.stream(s1):
            ds.copy_(hs, non_blocking=True)
        with torch.cuda.stream(s2):
            back.copy_(src, non_blocking=True)
    print("both directions ms", timeit(both))


ext = api.PlaneExtractor(max_frames=n, n_streams=st, enable_supposed=sup)
ext.set_upload_mode(mode)
F5 = float(np.float32(1.0) / np.float32(5000.0))
call = (lambda: ext.extract_batch_u16_ptr(host.data_ptr(), n, 480, 640, F5)) if U16 else (lambda: ext.extract_batch_ptr(host.data_ptr(), n, 480, 640))
for prof in (False, True):
    ext.set_profile(prof)
    for _ in range(3):
        call()
    t0 = time.perf_counter()
    call()
    print("profile", prof, "wall ms", (time.perf_counter() - t0) * 1e3, "xfer", ext.transfer_bytes())
tl = ext.kernel_timeline()
seen = {}
rows = {}
for name, a, b in tl:
    name = name.split("<")[0]
    g = seen.get(name, 0); seen[name] = g + 1
    rows.setdefault(g, []).append((name, a, b))
for g, r in sorted(rows.items()):
    first = min(a for _, a, _ in r); last = max(b for _, _, b in r)
    kb = [(a, b) for nm, a, b in r if nm == "k_border"]
    kl = [(a, b) for nm, a, b in r if nm == "k_lines"]
    print(f"g{g}: first kernel {first:7.3f}  lines {kl[0][0]:7.3f}->{kl[0][1]:7.3f}  border {kb[0][0]:7.3f}->{kb[0][1]:7.3f}  last kernel end {last:7.3f}" if kb else f"g{g}: {first:7.3f} -> {last:7.3f}")
if os.environ.get("SPX_TL_FULL"):
    for g in (range(st) if st <= 4 else (3, 7)):
        print(f"--- group {g}")
        for nm, a, b in sorted(rows.get(g, []) + [(n2 + "'", a, b) for n2, a, b in rows.get(g + st, [])], key=lambda r: r[1]):
            print(f"   {nm:18s} {a:7.3f} -> {b:7.3f} ({b - a:6.3f})")
