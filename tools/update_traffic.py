#!/usr/bin/env python
"""profiles/traffic.json from an ncu --set full capture (run here, no GPU needed), stamped with the hash of the CUDA
sources the capture was taken with: bench.py reports `roofline.traffic` only when the stamp matches the current sources.

usage: python tools/update_traffic.py <workload> <report.ncu-rep> <env_steps_per_launch>
       (repeat per workload; the stamp is refreshed every time, so re-capture every workload after a kernel change)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    workload, rep, env_steps = sys.argv[1], sys.argv[2], float(sys.argv[3])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    col = lambda name: hdr.index(name)
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        tj = json.load(open(path))
    except Exception:
        tj = {}
    stamp = bench.kernel_source_hash()
    if tj.get("kernel_source_hash") != stamp:   # a new build invalidates every earlier capture
        tj = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch and executed FP32 flops (2*FFMA + FADD + "
                          "FMUL thread instructions) per env-step, from ncu --set full captures of the build whose CUDA "
                          "sources hash to kernel_source_hash (tools/update_traffic.py)",
              "kernel_source_hash": stamp}
    entry = {"fp32_flops_per_env_step": {}}
    for r in rows:
        name = r[col("Kernel Name")]
        key = "rollout_forward_kernel" if "rollout_forward" in name else "rollout_backward_kernel" if "rollout_backward" in name else None
        if key is None:
            continue
        rd = float(r[col("dram__bytes_read.sum")]) * unit_scale[units[col("dram__bytes_read.sum")]]
        wr = float(r[col("dram__bytes_write.sum")]) * unit_scale[units[col("dram__bytes_write.sum")]]
        entry[key] = int(rd + wr)
        # the raw page holds these as per-cycle rates summed over the SMSPs: x elapsed SMSP cycles = thread instructions
        cyc = float(r[col("smsp__cycles_elapsed.avg")]) if "smsp__cycles_elapsed.avg" in hdr else float(r[col("sm__cycles_elapsed.avg")])
        f = lambda op: float(r[col("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op)]) * cyc
        flops = 2 * f("ffma") + f("fadd") + f("fmul")
        entry["fp32_flops_per_env_step"][key] = int(round(flops / env_steps))
    tj[workload] = entry
    json.dump(tj, open(path, "w"), indent=1)
    print(json.dumps(entry))


if __name__ == "__main__":
    main()
