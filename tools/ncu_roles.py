"""Per-role instruction budget of a kernel from an ncu source-page export:
  ncu -i X.ncu-rep --page source --csv --print-source sass,cuda --kernel-name regex:K > src.csv
  python tools/ncu_roles.py src.csv <main .cu file name> <tiles> line_lo:line_hi:name ...
SASS instructions are attributed to the role whose line range (in the main file) holds the nearest preceding instruction
of the main file in address order (inlined helpers from headers inherit the role of their call site)."""
import csv
import sys
from collections import defaultdict


def main():
    path, main_file, tiles = sys.argv[1], sys.argv[2], float(sys.argv[3])
    roles = []
    for a in sys.argv[4:]:
        lo, hi, name = a.split(":")
        roles.append((int(lo), int(hi), name))
    rows = list(csv.reader(open(path, errors="ignore")))
    sass = {}
    cur_file = None
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Function Name" or hdr is None:
            continue
        if r[0]:
            cur_line = int(r[0])
            continue
        addr = r[2]
        line = cur_line
        if addr.startswith("0x"):
            ie = hdr.index("Instructions Executed")
            try:
                n = float(r[ie])
            except ValueError:
                n = 0.0
            smp = float(r[hdr.index("# Samples")] or 0)
            sass[int(addr, 16)] = (cur_file, line, r[3].strip(), n, smp)
    by_role = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
    role = "prologue"
    for addr in sorted(sass):
        f, line, text, n, smp = sass[addr]
        if f.endswith(main_file):
            for lo, hi, name in roles:
                if lo <= line <= hi:
                    role = name
                    break
        by_role[role][0] += n
        by_role[role][1] += smp
        op = text.split()[0] if not text.startswith("@") else text.split()[1]
        by_role[role][2][op.split(".")[0]] += n
    tot = sum(v[0] for v in by_role.values())
    tot_s = sum(v[1] for v in by_role.values())
    print(f"total warp-instructions {tot:.0f} = {tot / tiles:.0f} per tile; samples {tot_s:.0f}")
    for name, (n, smp, ops) in sorted(by_role.items(), key=lambda kv: -kv[1][0]):
        top = ", ".join(f"{k} {v / tiles:.0f}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14])
        print(f"{name:12s} {n / tiles:8.0f} / tile  ({100 * n / tot:4.1f} % of instructions, {100 * smp / max(tot_s, 1):4.1f} % of samples)  {top}")


if __name__ == "__main__":
    main()
