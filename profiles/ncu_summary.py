#!/usr/bin/env python
"""Text summaries of ncu output for profiles/.

  python profiles/ncu_summary.py launches <launches.csv> [frames]       per-kernel totals of a `--metrics gpu__time_duration.sum` launch list
  python profiles/ncu_summary.py full <report.ncu-rep> [kernel regex]    key metrics of every profiled launch of a `--set full` capture
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def short(name):
    m = re.search(r"(k_\w+|Device\w+Kernel)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def launches(path, frames):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    ik, iv, im = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if r[im] == "gpu__time_duration.sum":
            tot[short(r[ik])] += float(r[iv].replace(",", "")) / 1e3
            cnt[short(r[ik])] += 1
    allus = sum(tot.values())
    print("# per-launch times are cold-cache and serialised (ncu): compare SHARES, not absolutes")
    for k, v in tot.most_common():
        print(f"{v / cnt[k]:9.1f} us avg  x {cnt[k]:4d}  {100 * v / allus:5.1f}%  {k}")
    print(f"# {sum(cnt.values())} launches, {allus:.1f} us in total" + (f" = {allus / frames:.1f} us and {sum(cnt.values()) / frames:.1f} launches per frame" if frames else ""))


def full(rep, kre):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        if kre and not re.search(kre, name):
            continue
        print(f"==== {short(name)}  grid {r[h.index('Grid Size')]} block {r[h.index('Block Size')]}")
        for k in KEYS:
            cols = [i for i, c in enumerate(h) if c == k or c.endswith("." + k)]
            if cols:
                print(f"  {k:75s} {r[cols[0]]} {units[cols[0]]}")
        stalls = [(float(r[i].replace(",", "")), c) for i, c in enumerate(h)
                  if "smsp__average_warps_issue_stalled" in c and c.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        for v, c in sorted(stalls, reverse=True)[:7]:
            print(f"  stall {c.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):40s} {v:.3f} warps/issue")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
