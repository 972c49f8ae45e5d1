"""Opcode mix of a kernel, overall and per source line, from an ncu report taken with --import-source on (development aid):
    python tools/ncu_opmix.py report.ncu-rep [kernel_regex] [top_lines]"""
import csv, re, collections, subprocess, sys, io
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + (["--kernel-name", "regex:" + kre] if kre else [])
rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
cur = None; line = None; per = collections.defaultdict(collections.Counter); ops = collections.Counter(); tot = 0; ii = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": ii = r.index("Instructions Executed"); continue
    if r[0].isdigit() and r[2] == "-": line = int(r[0]); continue
    if len(r) > 3 and r[2].startswith("0x"):
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3].strip())
        o = m.group(2) if m else r[3].strip(); n = int(r[ii])
        per[(cur, line)][o] += n; ops[o] += n; tot += n
over = ("BRA", "BSSY", "BSYNC", "ISETP", "IMAD", "IADD3", "VIADD", "LEA", "LOP3", "LDC", "LDCU", "MOV", "UMOV", "SHF", "SEL", "PRMT", "S2UR", "ULEA", "R2UR", "PLOP3", "UIADD3", "UISETP")
print("warp instructions", tot, "| overhead share %.1f%%" % (100.0 * sum(v for k, v in ops.items() if k in over) / max(tot, 1)))
print(" ".join(f"{a}:{b/tot*100:.1f}" for a, b in ops.most_common(32)))
lst = sorted(((sum(c.values()), sum(v for k, v in c.items() if k in over), k, c) for k, c in per.items()), reverse=True)
for t, o, k, c in lst[:top]:
    print(f"{k[0][:24]}:{k[1]:<4} {t/tot*100:4.1f}% (overhead {o/tot*100:4.1f}) | " + " ".join(f"{a}:{b/tot*100:.1f}" for a, b in c.most_common(8)))
