"""Per-source-line stall samples / instruction counts from an ncu report (needs -lineinfo and --import-source on):
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = ""; H = None; data = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": H = r; si = H.index("# Samples"); ii = H.index("Instructions Executed"); continue
    if H and r[0].isdigit() and r[2] == "-":
        try: data.append((int(r[si]), int(r[ii]), cur_file, int(r[0]), r[1].strip()[:100]))
        except ValueError: pass
ts = sum(d[0] for d in data) or 1; ti = sum(d[1] for d in data) or 1
print("total samples", ts, "warp instructions", ti)
for d in sorted(data, key=lambda d: -d[0])[:top]:
    print(f"{d[0]/ts*100:5.1f}% smp {d[1]/ti*100:5.1f}% inst  {d[2]}:{d[3]:<4} {d[4]}")
