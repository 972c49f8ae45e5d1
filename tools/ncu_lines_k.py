"""Per-source-line instruction / stall-sample shares of ONE kernel of an ncu report with several results:
    python tools/ncu_lines_k.py report.ncu-rep kernel_substring [min_pct]"""
import csv, subprocess, sys, io
rep, want = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + want],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H = None; data = []; cur = ""
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": H = r; si = H.index("# Samples"); ii = H.index("Instructions Executed"); continue
    if H and r[0].isdigit() and r[2] == "-":
        try: data.append((cur, int(r[0]), int(r[si]), int(r[ii]), r[1].strip()[:120]))
        except ValueError: pass
ti = sum(d[3] for d in data) or 1; ts = sum(d[2] for d in data) or 1
print("warp instructions", ti, "samples", ts)
for d in sorted(data, key=lambda d: (d[0], d[1])):
    if d[3] / ti * 100 >= minp or d[2] / ts * 100 >= minp:
        print(f"{d[0][:20]:20}:{d[1]:<4} inst {d[3]/ti*100:5.1f}%  smp {d[2]/ts*100:5.1f}%  {d[4]}")
