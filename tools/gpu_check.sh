#!/bin/bash
# GPU-box check used during development: parity tests, then the bench line with the per-kernel table.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 400 python bench.py --no-cpu-baseline "$@" > gpurun_out/b.json 2> gpurun_out/b.err; tail -3 gpurun_out/b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/b.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "path frac", round(d["roofline"]["path"]["frac"], 4))
for k in ("e2e", "e2e_whole_image", "e2e_u16"):
    if d.get(k):
        print(k, {a: b for a, b in d[k].items() if a != "note"})
print("lat", d["single_frame_latency_ms"])
print(d["roofline"]["kernels_ms"])
PY
