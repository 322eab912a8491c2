#!/usr/bin/env python
"""Kernel launches and GPU-busy time of one eager optimisation iteration of the reference-shaped imitation recipe
(laikago, mi-pace, 10 windows x 760 substeps), by kernel family.  Needs a GPU.  usage: python tools/count_iteration_launches.py"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppr_diffphys_b200.imitation import ImitationModel  # noqa: E402

torch.manual_seed(8)
m = ImitationModel("laikago", "mi-pace", total_iters=20, lr=1e-4)
m.record_forces = False
m.train()
m.reinit_envs(10, 24)
for _ in range(3):
    out = m(); m.backward(out["total_loss"]); m.update()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = m(); m.backward(out["total_loss"]); m.update()
    torch.cuda.synchronize()
fam = collections.Counter(); tim = collections.Counter()
for e in prof.events():
    if e.device_type is not None and str(e.device_type).endswith("CUDA") and e.device_time_total >= 0 and "Memcpy" not in e.name and "Memset" not in e.name:
        n = e.name
        k = ("rollout" if "rollout_" in n else "fk" if n.startswith("void fk_") or "fk_" in n.split("(")[0] else
             "se3_loss / frame_compose" if ("se3_loss" in n or "frame_compose" in n) else
             "gemm" if ("gemm" in n.lower() or "cutlass" in n.lower()) else "torch elementwise / reduce / copy")
        fam[k] += 1; tim[k] += e.device_time_total
tot = sum(fam.values())
print("kernel launches in one eager iteration: %d, GPU-busy %.3f ms" % (tot, sum(tim.values()) / 1e3))
for k, v in fam.most_common():
    print("  %-36s %5d launches  %8.3f ms" % (k, v, tim[k] / 1e3))
if "--names" in sys.argv:
    byname = collections.Counter(); tn = collections.Counter()
    for e in prof.events():
        if e.device_type is not None and str(e.device_type).endswith("CUDA") and e.device_time_total >= 0:
            byname[e.name[:150]] += 1; tn[e.name[:150]] += e.device_time_total
    for k, v in byname.most_common():
        print("  %3d x %7.1f us  %s" % (v, tn[k] / v, k))
