#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): per-kernel headline metrics, opcode mix, stall reasons.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [warp_steps_per_launch]"""
import collections
import csv
import io
import re
import subprocess
import sys


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    wsteps = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
            "smsp__cycles_active.avg", "sm__cycles_elapsed.avg"]
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("%-70s %-16s %s" % (w, units[i], [r[i][:60] for r in rows]))
    if wsteps and "smsp__inst_executed.sum" in hdr:
        i = hdr.index("smsp__inst_executed.sum")
        print("warp instructions per warp-substep:", [round(float(r[i]) / wsteps, 1) for r in rows])
    names = [r[hdr.index("Kernel Name")] for r in rows]
    for kn in dict.fromkeys(names):
        short = re.match(r"[\w:]+", kn.replace("void ", "")).group(0)
        src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--kernel-name",
                                                "regex:" + short.split("::")[-1]]))))
        his = [i for i, r in enumerate(src) if "Instructions Executed" in r]
        if not his:
            continue
        h = src[his[0]]
        iI, iS = h.index("Instructions Executed"), h.index("Source")
        stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        ops, stalls, tot = collections.Counter(), collections.Counter(), 0.0
        for r in src[his[0] + 1:]:
            if len(r) != len(h):
                continue
            try:
                ie = float(r[iI] or 0)
            except ValueError:
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
            ops[m.group(2).split(".")[0] if m else "?"] += ie
            tot += ie
            for i in stall_cols:
                stalls[h[i]] += float(r[i] or 0)
        print("== %s: static SASS %d" % (short, len(src) - his[0] - 1))
        print("   opcodes: " + "  ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in ops.most_common(20)))
        ts = sum(stalls.values()) or 1
        print("   stalls : " + "  ".join("%s %.1f%%" % (k[6:], 100 * v / ts) for k, v in stalls.most_common(9)))


if __name__ == "__main__":
    main()
