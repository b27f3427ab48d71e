"""Warp-instructions executed per source line (per unit) from an ncu source-page export:
  python tools/ncu_lines.py src.csv <file name> <units> [line_lo line_hi]   (inlined header code is charged to the call-site line of the main file)"""
import csv, sys
from collections import defaultdict
path, main_file, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 10 ** 9)
rows = list(csv.reader(open(path, errors="ignore")))
hdr = cur = line = None
sass = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0]: line = int(r[0]); continue
    if r[2].startswith("0x"):
        d = dict(zip(hdr, r))
        sass.append((int(r[2], 16), cur.split("/")[-1], line, r[3].strip(), float(d["Instructions Executed"] or 0), float(d["# Samples"] or 0)))
sass.sort()
per = defaultdict(lambda: [0.0, 0.0, 0, defaultdict(float)])
site = 0
for a, f, l, t, n, s in sass:
    if f == main_file: site = l
    p = per[site]; p[0] += n; p[1] += s; p[2] += 1
    op = (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
    p[3][op] += n
src = open([x for x in __import__("glob").glob("reformer_tts_b200/csrc/" + main_file)][0]).read().split("\n")
for l in sorted(per):
    if lo <= l <= hi and per[l][0] / units >= 1.0:
        ops = ", ".join(f"{k} {v / units:.0f}" for k, v in sorted(per[l][3].items(), key=lambda kv: -kv[1])[:6])
        print(f"{l:4d} {per[l][0] / units:7.0f} instr  {per[l][1]:5.0f} smp  {per[l][2]:4d} sass | {src[l - 1].strip()[:70]:70s} | {ops}")
