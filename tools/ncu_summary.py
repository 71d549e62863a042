"""Summarise ncu output into the small text/JSON files kept under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.md>
  python tools/ncu_summary.py full <report.ncu-rep> <out.md> [traffic.json kernel_regex]
"""
import csv
import json
import re
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hi]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    tot = defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) > vi:
            name = r[ki].split("(")[0]
            if name.startswith("void "):            # template instantiations print their return type
                name = name[5:]
            tot[name].append(float(r[vi].replace(",", "")))
    ours = {k: v for k, v in tot.items() if k.startswith("dvc::")}
    s = sum(sum(v) for v in ours.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | avg us (cold, serialised) | share of dvc:: time |\n|---|---|---|---|\n")
        for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1000:.1f} | {sum(v) / s * 100:.1f}% |\n")
        other = sum(sum(v) for k, v in tot.items() if not k.startswith("dvc::"))
        f.write(f"\nnon-dvc launches in the capture (input generation by torch, outside the timed "
                f"region): {other / 1e6:.2f} ms total\n")
    print(open(out).read())


def full(rep, out, traffic_json=None, kernel_re=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    with open(out, "w") as f:
        for r in rows[2:]:
            name = r[ki].split("(")[0]
            f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---|---|\n")
            vals = {}
            for k in KEYS:
                if k in h:
                    i = h.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
                    vals[k] = (r[i], units[i])
            f.write("\n")
            if traffic_json and kernel_re and re.search(kernel_re, name):
                def to_bytes(key):
                    v, u = vals[key]
                    v = float(v.replace(",", ""))
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                tb = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
                json.dump({"kernel": name, "dram_bytes_per_launch": tb,
                           "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
                           "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
                           "source": rep.split("/")[-1]}, open(traffic_json, "w"), indent=1)
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(*sys.argv[2:])
