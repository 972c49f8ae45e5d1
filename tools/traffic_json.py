"""profiles/r2_traffic_all.json from an ncu metrics pass over a resident step run as ONE frame group:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/traffic.csv python tools/prof_run.py 1000 1
    python tools/traffic_json.py gpurun_out/traffic.csv 1000 profiles/r2_traffic_all.json
The last of the three passes prof_run.py makes is used (every kernel once per pass)."""
import csv, json, sys

src, frames, dst = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iu, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
launches = {}
for r in rows:
    if r is hdr or not r[iid].isdigit():
        continue
    L = launches.setdefault(int(r[iid]), {"kernel": r[ik].split("(")[0].split("<")[0].replace("spx::", "").replace("void ", "").strip()})
    L[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
ids = sorted(launches)
per_pass = len(ids) // 3
last = [launches[i] for i in ids[-per_pass:]]
kern = {}
for L in last:
    k = kern.setdefault(L["kernel"], {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "us": 0.0})
    k["launches"] += 1
    k["dram_read_bytes"] += L.get("dram__bytes_read.sum", 0.0)
    k["dram_write_bytes"] += L.get("dram__bytes_write.sum", 0.0)
    k["us"] += L.get("gpu__time_duration.sum", 0.0)
for k in kern.values():
    k["dram_bytes_per_frame"] = (k["dram_read_bytes"] + k["dram_write_bytes"]) / frames
tot = sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in kern.values())
out = {"frames": frames, "launches_in_pass": per_pass, "path_dram_bytes_per_frame": tot / frames,
       "path_us_serialised": sum(k["us"] for k in kern.values()),
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over python tools/prof_run.py "
              f"{frames} 1 (one frame group: every kernel once per pass); last of three passes; memsets and copies are not kernels and are not counted",
       "kernels": dict(sorted(kern.items(), key=lambda kv: -kv[1]["dram_bytes_per_frame"]))}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps({"path_dram_bytes_per_frame": out["path_dram_bytes_per_frame"], "launches": per_pass}))
