#!/usr/bin/env python
"""Motion imitation on the B200 rollout path -- the reference's `run.sh` recipe
(`python main.py --urdf_template laikago --seqname mi-pace --logname 0`, /root/reference/run.sh:12,
/root/reference/main.py:50-105) on top of ppr_diffphys_b200.

    python examples/run_imitation.py --seqname mi-pace --iters 101 [--num-envs 10 --frames-per-wdw 24]

Prints the loss every `--log-every` iterations and a JSON summary (iteration time, env-steps/s)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

from ppr_diffphys_b200.imitation import GraphedStep, ImitationModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robot", default="laikago")
    ap.add_argument("--seqname", default="mi-pace")
    ap.add_argument("--iters", type=int, default=101)
    ap.add_argument("--num-envs", type=int, default=10)
    ap.add_argument("--frames-per-wdw", type=int, default=24)
    ap.add_argument("--log-every", type=int, default=10)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--no-graph", action="store_true", help="eager iterations instead of one CUDA-graph replay each")
    ap.add_argument("--eager-update", action="store_true",
                    help="capture forward + backward only; gradient norms, clipping and AdamW run eagerly after each replay")
    args = ap.parse_args()
    torch.manual_seed(8)
    model = ImitationModel(args.robot, args.seqname, total_iters=args.iters, lr=args.lr)
    model.record_forces = False
    model.train()
    model.reinit_envs(args.num_envs, args.frames_per_wdw)
    T = len(model.steps_idx)
    losses, times = [], []
    step = None if args.no_graph else GraphedStep(model, capture_update=not args.eager_update)
    for it in range(args.iters):
        model.progress = it / max(1, args.iters - 1)
        if it % 20 == 0:
            model.save_checkpoint()     # main.py:73-74: every iters_per_round iterations (in-memory roll-back queue)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if step is not None:
            loss_dict, info = step()
        else:
            loss_dict = model()
            model.backward(loss_dict["total_loss"])
            info = model.update()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        losses.append(float(loss_dict["loss_traj"].detach()))
        if it % args.log_every == 0 or it == args.iters - 1:
            print("it %4d  loss_traj %.5f  total %.6f  |grad| %.3f  %.1f ms" %
                  (it, losses[-1], float(loss_dict["total_loss"].detach()), info["grad_norm"], times[-1] * 1e3), flush=True)
    k = max(1, len(losses) // 10)
    steady = sorted(times[len(times) // 5:])
    med = steady[len(steady) // 2]
    print(json.dumps({"robot": args.robot, "seqname": args.seqname, "iters": args.iters, "num_envs": args.num_envs,
                      "substeps": T, "loss_traj_first": sum(losses[:k]) / k, "loss_traj_last": sum(losses[-k:]) / k,
                      "cuda_graph": step is not None, "median_iter_ms": med * 1e3, "env_steps_per_sec": args.num_envs * (T - 1) / med}))


if __name__ == "__main__":
    main()
