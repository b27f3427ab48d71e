"""SASS listing with executed counts, samples and the top stall reasons per instruction from an ncu source-page export:
  python tools/ncu_sass.py src.csv <addr_lo hex> <addr_hi hex>"""
import csv, sys
path, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
rows = list(csv.reader(open(path, errors="ignore")))
hdr = cur = line = None
out = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0]: line = int(r[0]); continue
    if r[2].startswith("0x"):
        a = int(r[2], 16)
        if lo <= a <= hi:
            d = dict(zip(hdr, r))
            st = sorted(((float(v), k.replace("stall_", "")) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "-", "0")), reverse=True)[:3]
            out[a] = f"{a:06x} {cur.split('/')[-1][:18]:18s}:{line:4d} {float(d['Instructions Executed'] or 0) / 1e3:8.0f}k {d['# Samples']:>5s}  {r[3].strip()[:70]:70s} " + " ".join(f"{k}={int(v)}" for v, k in st)
for a in sorted(out): print(out[a])
