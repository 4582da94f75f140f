#!/usr/bin/env python3
"""Reads one `ncu --set full --import-source on` report (gpurun_out/*.ncu-rep) and records, for every kernel in it,
what bench.py and DESIGN.md cite:

  profiles/ncu_<tag>.md           duration, registers, occupancy, issue-slot / FMA-pipe / LSU utilisation, DRAM bytes, stall mix,
                                  executed instruction mix and EXECUTED flop (FFMA2 = 4, FFMA = 2, FADD2 / FMUL2 = 2, FADD / FMUL = 1 per thread)
  profiles/kernel_traffic.json    dram__bytes_read.sum + dram__bytes_write.sum per launch, scaled to 1024 streams   (key = --key)
  profiles/kernel_exec.json       executed flop per launch (same scaling), FMA-pipe-active and issue-active percentages

usage: ncu_kernel_report.py <report.ncu-rep> <tag> --streams N [--key cascade0|channelize|audio|waterfall|...] [--note "..."]
"""
import argparse
import csv
import json
import os
import subprocess
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit"),
    ("lts__t_sector_hit_rate.pct", "L2 hit"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
]
FLOP = {"FFMA2": 4, "FFMA": 2, "FADD2": 2, "FMUL2": 2, "FADD": 1, "FMUL": 1}


def tobytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("tag")
    ap.add_argument("--streams", type=int, required=True)
    ap.add_argument("--key", default=None)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    src = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    src_rows = list(csv.reader(src.splitlines()))
    out = ["# ncu %s" % a.tag, "", "`ncu --set full --clock-control none --import-source on`, one launch, %d streams (times under the profiler are cold-cache"
           " and serialised: compare shares and ratios, not absolutes).  %s" % (a.streams, a.note), ""]
    scale = 1024.0 / a.streams
    traffic, execd = {}, {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        out += ["## `%s`" % name, "", "| metric | value |", "|---|---|"]
        for k, label in KEYS:
            if k in idx:
                out.append("| %s | %s %s |" % (label, r[idx[k]], units[idx[k]]))
        stalls = sorted(((float(r[i] or 0), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")), reverse=True)
        out.append("| stalls (warps per issue) | %s |" % ", ".join("%s %.2f" % (h.split("issue_stalled_")[1].split("_per_issue")[0], v) for v, h in stalls[:6] if v > 0.02))
        out.append("")
        try:
            tr = tobytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + tobytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            traffic[name] = tr
        except Exception:
            tr = None
        # executed instruction mix from the source page (first kernel of the page only matches single-kernel reports)
        if len(src_rows) > 2 and len(rows) == 3:
            sh = src_rows[1]
            si = {h: i for i, h in enumerate(sh)}
            agg, thr = defaultdict(float), defaultdict(float)
            for q in src_rows[2:]:
                s = q[si["Source"]].strip()
                if not s:
                    continue
                op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
                try:
                    agg[op] += float(q[si["Instructions Executed"]])
                    thr[op] += float(q[si["Predicated-On Thread Instructions Executed"]])
                except (ValueError, KeyError):
                    pass
            tot = sum(agg.values())
            flop = sum(thr[o] * f for o, f in FLOP.items())
            out.append("Executed warp instructions: %.4g; mix: %s." % (tot, ", ".join("%s %.1f %%" % (o, 100 * v / tot) for o, v in sorted(agg.items(), key=lambda x: -x[1])[:12])))
            out.append("")
            out.append("Executed FP32 work: **%.4g flop per launch** at %d streams (FFMA2 = 4 flop per thread, FFMA = 2, FADD2 / FMUL2 = 2, FADD / FMUL = 1)." % (flop, a.streams))
            out.append("")
            execd[name] = {"executed_flop_per_launch": flop * scale, "warp_instructions": tot * scale,
                           "fma_pipe_active_pct": float(r[idx["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]]),
                           "issue_active_pct": float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                           "source": "%s (%d streams, scaled to 1024)" % (a.tag, a.streams)}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "ncu_%s.md" % a.tag), "w") as f:
        f.write("\n".join(out) + "\n")
    if a.key:
        for fn, data in (("kernel_traffic.json", {k: v * scale for k, v in traffic.items()}), ("kernel_exec.json", execd)):
            path = os.path.join(ROOT, "profiles", fn)
            cur = {}
            if os.path.exists(path):
                with open(path) as f:
                    cur = json.load(f)
            if data:
                cur[a.key] = list(data.values())[0]
                if fn == "kernel_traffic.json":
                    cur.setdefault("_sources", {})[a.key] = "%s: one launch at %d streams, scaled to 1024" % (a.tag, a.streams)
                with open(path, "w") as f:
                    json.dump(cur, f, indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main()
