#!/usr/bin/env python
"""Generates tests/golden/rollout_<robot>.npz with the float64 oracle (oracle/sim_oracle.py).

The reference ships no golden vectors and cannot run here (warp_lang not installable), so these fixtures pin the
ORACLE's outputs (parity unpinned w.r.t. Warp itself): seeded synthetic inputs rounded to float32, oracle forward
(pos, vel, grf, jaf) and the 10 gradients of a fixed quadratic loss.  Re-run after changing the oracle:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from helpers import make_inputs, settle_height  # noqa: E402
from oracle import sim_oracle as so  # noqa: E402
from ppr_diffphys_b200 import load_robot  # noqa: E402

DT, STRIDE, F = 5e-4, 32, 3
KEYS = ["q_init", "qd_init", "torques", "res_f", "refs", "target_ke", "target_kd", "body_inv_mass", "body_inertia",
        "body_inv_inertia"]


def loss_weights(shape_pos, shape_vel, seed=11):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape_pos, generator=g, dtype=torch.float64),
            0.1 * torch.randn(shape_vel, generator=g, dtype=torch.float64))


def run(robot, bs, seed, margin, tag=None):
    rm = load_robot(robot)
    T = STRIDE * (F - 1) + 1
    rm, d = make_inputs(rm, bs=bs, T=T, seed=seed, lin_vel=0.5, res_f_std=0.05, torque_std=0.05, ang=0.25)
    # keep joint angles away from 0 (the literal acos twist angle of the reference is singular there)
    ja = d["q_init"][:, 7:]
    d["q_init"][:, 7:] = torch.where(ja.abs() < 0.05, 0.05 * torch.sign(ja) + (ja == 0) * 0.05, ja)
    if margin is not None:
        d = settle_height(rm, d, penetration=-margin)
    else:
        d["q_init"][:, 1] = 1.0  # airborne: no ground contact during the window
    d = {k: v.float().double() for k, v in d.items()}  # float32-representable inputs
    m = so.OracleModel(rm)
    a = {k: d[k].clone().requires_grad_(True) for k in KEYS}
    pos, vel, grf, jaf = so.rollout(m, a["q_init"], a["qd_init"], a["torques"], a["res_f"], a["refs"], a["target_ke"],
                                    a["target_kd"], a["body_inv_mass"], a["body_inertia"], a["body_inv_inertia"], DT,
                                    STRIDE, F)
    wp, wv = loss_weights(pos.shape, vel.shape)
    loss = 0.5 * ((pos - wp * 0.01) ** 2 * 10).sum() + (vel * wv).sum()
    grads = torch.autograd.grad(loss, [a[k] for k in KEYS])
    out = {"in_" + k: d[k].numpy().astype(np.float32) for k in KEYS}
    out.update(pos=pos.detach().numpy(), vel=vel.detach().numpy(), grf=grf.numpy(), jaf=jaf.numpy(),
               adj_pos=(10 * (pos - wp * 0.01)).detach().numpy(), adj_vel=wv.numpy(), loss=float(loss))
    out.update({"grad_" + k: g.numpy() for k, g in zip(KEYS, grads)})
    out.update(dt=DT, stride=STRIDE, nframes=F, robot=robot)
    path = os.path.join(HERE, "rollout_%s.npz" % (tag or robot))
    np.savez_compressed(path, **out)
    print(robot, "bs", bs, "max grf", float(grf.abs().max()), "loss", float(loss), "->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    run("laikago", 3, 21, -0.002)
    run("laikago", 3, 24, None, tag="laikago_air")
    run("human", 3, 22, -0.003)
    run("quad", 2, 23, -0.003)
