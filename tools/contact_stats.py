#!/usr/bin/env python
"""Histogram of ACTIVE contact points per body and substep in a bench workload, read back from the checkpoint
records the forward kernel writes (low halfword of float 2 of quad 3 of each row = count).  Needs a GPU.
usage: python tools/contact_stats.py [workload] [envs]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ppr_diffphys_b200 import SimEnv, load_robot  # noqa: E402
from ppr_diffphys_b200.synth import make_batch  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
    w = bench.WORKLOADS[wl]
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    T = w["window"] + 1
    rm = load_robot(w["robot"])
    env = SimEnv(rm)
    env.set_latency_envs(0)
    dev = env.device
    b = make_batch(env, bs, T, seed=0, clearance=w["clearance"], lin_vel=w["lin_vel"], pinned_host=False)
    t = lambda x: torch.as_tensor(x, device=dev)
    m = t(rm.body_mass)
    nI = t(rm.norm_body_inertia)
    I = nI * m[:, None, None]
    pos, vel, grf, jaf, ws = env.rollout_forward(bs, T, w["stride"], bench.DT, b["q_init"].to(dev), b["qd_init"].to(dev),
                                                 None, None, b["refs"].to(dev), t(rm.joint_target_ke),
                                                 t(rm.joint_target_kd), 1 / m, I, torch.linalg.inv(I),
                                                 want_forces=False, shared_params=True)
    torch.cuda.synchronize()
    threads, epg = env.packing
    ngroups = -(-bs // epg)
    nwarps = ngroups * (threads // 32)
    rq = 6 if all(int(t) in (1, 4) for t in rm.joint_type) else 7                 # float4 quads per body and row
    rows = ws[: (T - 1) * nwarps * rq * 128].view(T - 1, nwarps, rq, 32, 4)
    cnt_all = (rows[:, :, 3, :, 2].contiguous().view(torch.int32) & 0xffff).cpu().numpy()   # T-1, nwarps, 32
    cnt = cnt_all
    cnt = cnt.reshape(T - 1, ngroups, threads)[:, :, : epg * rm.nb]
    if threads > 32:   # block layout: slot = body * epg + env_in_group
        cnt = cnt.reshape(T - 1, ngroups, rm.nb, epg).transpose(0, 1, 3, 2)
    else:              # warp layout: slot = env_in_group * nb + body
        cnt = cnt.reshape(T - 1, ngroups, epg, rm.nb)
    print("workload %s, %d envs, packing %d threads / %d envs" % (wl, bs, threads, epg))
    print("mean active points per env-substep: %.2f" % cnt.sum(-1).mean())
    # block layout: columns are POSITIONS in the block (the library orders the bodies to balance the contact-prone
    # leaf meshes over the warps), not body indices
    print("per body (block layout: per position) mean:", np.round(cnt.mean((0, 1, 2)), 2))
    print("per body max :", cnt.max((0, 1, 2)))
    nz = cnt[cnt > 0]
    print("bodies with contact per env-substep: %.2f; points per touching body: mean %.2f, p50 %d, p90 %d, p99 %d, max %d"
          % ((cnt > 0).sum(-1).mean(), nz.mean(), *np.percentile(nz, [50, 90, 99]).astype(int), nz.max()))
    # the serial cost of the owner loop per warp = max over the warp's lanes; flattened = ceil(sum / 32)
    lanes = cnt_all.clip(0, 9)
    print("per warp-substep: max over lanes %.2f (serial owner loop trips), sum over lanes %.2f (flattened work items)"
          % (lanes.max(-1).mean(), lanes.sum(-1).mean()))
    print("by time: ", np.round(cnt.sum(-1).mean((1, 2))[:: max(1, (T - 1) // 16)], 1))


if __name__ == "__main__":
    main()
