#!/usr/bin/env python3
"""Turns an ncu report (gpurun_out/*.ncu-rep, --set full) and a launch list CSV into the tracked summaries under
profiles/.  usage: summarize_ncu.py <report.ncu-rep> <launches.csv> <tag>"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
]


def main():
    rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    streams = int(sys.argv[4]) if len(sys.argv) > 4 else 256
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = ["# ncu summary %s" % tag, "", "Source: `ncu --set full --clock-control none --import-source on` on a B200 (one launch per kernel,",
           "%d streams x 2.4 Msps x 1 s; absolute times are cold-cache/serialised -- compare shares)." % streams, ""]
    traffic = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        out.append("## `%s`" % name)
        out.append("")
        out.append("| metric | value |")
        out.append("|---|---|")
        for k, label in KEYS:
            if k in idx:
                out.append("| %s | %s %s |" % (label, r[idx[k]], units[idx[k]]))
        out.append("")
        try:
            def tob(v, u):
                v = float(v.replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            traffic[name] = tob(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                tob(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        except Exception:
            pass
    per = OrderedDict()
    if os.path.exists(launches):
        with open(launches) as f:
            lines = [ln for ln in f if not ln.startswith("==")]
        for row in csv.DictReader(lines):
            if row.get("Metric Name") == "gpu__time_duration.sum":
                k = row["Kernel Name"].split("(")[0]
                v = float(row["Metric Value"].replace(",", ""))
                v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row.get("Metric Unit", "ns"), 1e-6)
                per.setdefault(k, []).append(v)
        tot = sum(sum(v) for v in per.values())
        out += ["## launch list of `python bench.py --steps 2 --warmup 1` (gpu__time_duration.sum)", "",
                "| kernel | launches | total ms | share |", "|---|---|---|---|"]
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            out.append("| `%s` | %d | %.3f | %.1f %% |" % (k[:90], len(v), sum(v), 100 * sum(v) / tot))
        out.append("")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "ncu_%s_summary.md" % tag), "w") as f:
        f.write("\n".join(out))
    # per-launch DRAM traffic keyed like bench.py's kernel table, scaled to 1024 streams
    keymap = (("cascade_kernel<0", "cascade0"), ("hbarb_tile", "cascade1"), ("cascade_kernel<2", "cascade1"), ("channelize16", "channelize"),
              ("audio_fft", "audio"), ("audio_kernel", "audio"))
    kt = {}
    for name, v in traffic.items():
        for pat, key in keymap:
            if pat in name and key not in kt:
                kt[key] = v * 1024 / streams
    if kt:
        kt["_source"] = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch at %d streams, scaled to 1024 (%s)" % (streams, tag)
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json"), "w") as f:
            json.dump(kt, f, indent=1)
    print("\n".join(out[-14:]))


if __name__ == "__main__":
    main()
