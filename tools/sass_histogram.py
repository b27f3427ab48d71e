"""Per-kernel SASS opcode histogram of libreformer_b200.so (what proves a Blackwell-native kernel: UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA loads, LDGSTS = cp.async, HMMA would be the legacy mma.sync path).

    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "reformer_tts_b200", "libreformer_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "SYNCS", "FFMA2", "HSET2", "MUFU.EX2", "REDG", "RED.", "ATOMG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    kernels = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = demangle(m.group(1))
            kernels[name] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            kernels[name]["_total"] += 1
            for k in KEYS:
                if m.group(1).startswith(k):
                    kernels[name][k] += 1
    from reformer_tts_b200.csrc.build import source_hash
    print(f"# cuobjdump -sass reformer_tts_b200/libreformer_b200.so   (build id {source_hash()}; static instruction counts per kernel)")
    print(f"{'kernel':96s} {'instr':>6s} " + " ".join(f"{k:>8s}" for k in KEYS))
    tot = collections.Counter()
    for name, c in kernels.items():
        short = re.sub(r"\(.*", "", name).replace("void ", "")
        print(f"{short[:96]:96s} {c['_total']:6d} " + " ".join(f"{c[k]:8d}" for k in KEYS))
        tot.update(c)
    print(f"{'TOTAL':96s} {tot['_total']:6d} " + " ".join(f"{tot[k]:8d}" for k in KEYS))


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
