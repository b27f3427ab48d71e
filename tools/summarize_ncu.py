"""Summarise ncu CSV output into small text tables for profiles/.
  python tools/summarize_ncu.py launches <launches.csv>      per-kernel count / total time / share (gpu__time_duration pass)
  python tools/summarize_ncu.py full <raw.csv>               key metrics of each profiled launch (--set full, --page raw --csv)"""
import csv
import re
import sys
from collections import defaultdict

FULL_METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
                "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
                "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
                "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
                "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)[:90]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    data = rows[rows.index(hdr) + 1:]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in data:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)     # -> microseconds
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print(f"{sum(a[0] for a in agg.values())} launches, {total / 1e3:.3f} ms summed kernel time (ncu: cold cache, serialised - compare SHARES, not absolutes)")
    own = ("rtts::", "f64p::", "f64::", "xa::")      # (ncu prints the kernels of nested namespaces without the outer rtts::)
    ours = sum(a[1] for k, a in agg.items() if k.startswith(own))
    spin = sum(a[1] for k, a in agg.items() if "spin_kernel" in k)
    print(f"libreformer_b200 kernels (rtts::*, incl. f64p:: / xa::): {100 * ours / total:.1f} % of kernel time, {100 * ours / max(total - spin, 1e-9):.1f} % without the "
          f"kernel timer's spin kernels")
    print(f"{'kernel':92s} {'n':>6s} {'total us':>11s} {'avg us':>9s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{k:92s} {n:6d} {t:11.1f} {t / n:9.1f} {100 * t / total:6.1f}%")


def full(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", short(r[hdr.index("Kernel Name")]))
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:88s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
