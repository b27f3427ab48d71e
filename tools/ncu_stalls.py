"""Stall reasons per role (work instructions vs barrier-wait loops) from an ncu source-page export, and the hottest SASS lines:
  python tools/ncu_stalls.py src.csv <main .cu> <units> <wait line lo:hi> lo:hi:name ..."""
import csv
import sys
from collections import defaultdict

path, main_file, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
wlo, whi = map(int, sys.argv[4].split(":"))
roles = [(int(a.split(":")[0]), int(a.split(":")[1]), a.split(":")[2]) for a in sys.argv[5:]]
rows = list(csv.reader(open(path, errors="ignore")))
hdr = cur = line = None
sass = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0]:
        line = int(r[0])
        continue
    if r[2].startswith("0x"):
        sass.append((int(r[2], 16), cur.split("/")[-1], line, r[3].strip(), dict(zip(hdr, r))))
sass.sort()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
waitops = ("SYNCS", "NOP", "YIELD", "BRA", "VIADD", "ISETP", "BPT", "UMOV", "WARPSYNC", "BSSY", "BSYNC", "NANOSLEEP")
role = "prologue"
agg = defaultdict(lambda: defaultdict(float))
static = defaultdict(int)
for a, f, l, t, d in sass:
    if f == main_file:
        for lo, hi, n in roles:
            if lo <= l <= hi:
                role = n
    op = (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
    is_wait = (f != main_file and op in waitops and f in ("common.cuh", "sm_30_intrinsics.hpp")) or (f == main_file and wlo <= l <= whi)
    key = role + (":wait" if is_wait else ":work")
    static[role] += 1
    for k in stall_cols:
        v = d[k]
        if v not in ("", "-"):
            agg[key][k] += float(v)
    agg[key]["n"] += float(d["Instructions Executed"] or 0)
print("static SASS instructions per role:", dict(static), "total", sum(static.values()), "=", sum(static.values()) * 16 // 1024, "KB")
for k, v in sorted(agg.items()):
    tot = sum(x for kk, x in v.items() if kk != "n")
    print(f"{k:16s} instr/unit={v['n'] / units:7.0f} samples={tot:6.0f}", {kk.replace("stall_", ""): int(x) for kk, x in sorted(v.items(), key=lambda z: -z[1]) if kk != "n" and x > tot * 0.04})
