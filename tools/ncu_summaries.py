#!/usr/bin/env python
"""Turn ncu exports into the markdown summaries kept under profiles/.

  python tools/ncu_summaries.py launches <launch-list.csv> <out.md> "<title>" "<command>"
  python tools/ncu_summaries.py full <raw.csv from `ncu -i X.ncu-rep --page raw --csv`> <out.md> "<title>" "<command>" [traffic.json pairs_per_launch]
"""
import collections
import csv
import json
import re
import sys


def launches(path, out, title, command):
    rows = list(csv.reader(open(path)))
    hdr, agg, total, n = None, collections.OrderedDict(), 0.0, 0
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
                ki, vi = r.index("Kernel Name"), r.index("Metric Value")
            continue
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("gtb::", "")
        name = re.sub(r"RsCfg<[^>]*>", "Cfg", name)[:110]
        v = float(r[vi].replace(",", "")) / 1e6
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    with open(out, "w") as fh:
        fh.write(f"# {title}\n\nCommand: `{command}`\n(per-launch times are cold-cache and serialised: compare shares)\n\n")
        fh.write("| launches | total ms | share | kernel |\n|---:|---:|---:|---|\n")
        for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            fh.write(f"| {c} | {v:.3f} | {100 * v / total:.1f}% | `{k}` |\n")
        os_ = sum(v for k, (c, v) in agg.items() if "rs_onesweep_kernel" in k)
        fh.write(f"\nTotal {total:.3f} ms over {n} launches. The onesweep passes are {100 * os_ / total:.1f} % of the device time.\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "lts__t_sector_hit_rate.pct"]


def full(path, out, title, command, traffic=None, pairs=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    data = rows[2:]
    with open(out, "w") as fh:
        fh.write(f"# {title}\n\nCommand: `{command}`\n\n")
        fh.write("| metric | unit | " + " | ".join(f"launch {i + 1}" for i in range(len(data))) + " |\n")
        fh.write("|---|---|" + "---:|" * len(data) + "\n")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                fh.write(f"| {w} | {units[i]} | " + " | ".join(r[i] for r in data) + " |\n")
        fh.write("\nKernels: " + "; ".join("`" + r[hdr.index("Kernel Name")][:90] + "`" for r in data) + "\n")
        if traffic and pairs:
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            # (launches far shorter than the longest are another instantiation on a tiny input: not averaged)
            it = hdr.index("gpu__time_duration.sum")
            tmax = max(float(r[it].replace(",", "")) for r in data)
            data = [r for r in data if float(r[it].replace(",", "")) >= 0.1 * tmax]

            def gb(r, i):
                v = float(r[i].replace(",", ""))
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
            per = sum(gb(r, ir) + gb(r, iw) for r in data) / len(data)
            inst = sum(float(r[hdr.index("smsp__inst_executed.sum")].replace(",", "")) for r in data) / len(data)
            fh.write(f"\nDRAM traffic per launch = {per / 1e9:.3f} GB for {pairs:.3g} pairs = {per / pairs:.2f} B/pair; "
                     f"algorithmic 24 B/pair (12 read + 12 written).\n"
                     f"Warp instructions per launch = {inst / 1e6:.0f} M = {inst / pairs * 32:.0f} per 32 pairs.\n")
            json.dump({"kernel": data[0][hdr.index("Kernel Name")][:120], "dram_bytes_per_launch": per,
                       "pairs_per_launch": pairs, "dram_bytes_per_pair": per / pairs, "source": out}, open(traffic, "w"),
                      indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:6])
    else:
        a = sys.argv[2:]
        full(a[0], a[1], a[2], a[3], a[4] if len(a) > 4 else None, float(a[5]) if len(a) > 5 else None)
