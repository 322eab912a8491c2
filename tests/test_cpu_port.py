"""The CPU port (oracle/cpu_port: the timed CPU baseline, sharing the per-body arithmetic header with the CUDA
kernels) against the independent float64 autograd oracle and the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle.cpu_port import CpuRollout
from ppr_diffphys_b200 import load_robot

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEYS = ["q_init", "qd_init", "torques", "res_f", "refs", "target_ke", "target_kd", "body_inv_mass", "body_inertia",
        "body_inv_inertia"]


@pytest.mark.parametrize("fixture", ["laikago", "laikago_air", "human", "quad"])
def test_cpu_port_f64_matches_golden(fixture):
    z = np.load(os.path.join(GOLDEN, "rollout_%s.npz" % fixture))
    rm = load_robot(str(z["robot"]))
    d = {k: torch.from_numpy(z["in_" + k]).double() for k in KEYS}
    cpu = CpuRollout(rm)
    pos, vel, grf, jaf = cpu.forward(d, float(z["dt"]), int(z["stride"]), int(z["nframes"]), want_forces=True)
    assert (pos - torch.from_numpy(z["pos"])).abs().max() < 1e-10
    assert (vel - torch.from_numpy(z["vel"])).abs().max() < 1e-9
    assert (grf - torch.from_numpy(z["grf"])).abs().max() < 1e-7
    assert (jaf - torch.from_numpy(z["jaf"])).abs().max() < 1e-7
    g = cpu.backward(torch.from_numpy(z["adj_pos"]), torch.from_numpy(z["adj_vel"]))
    for k in KEYS:
        ref = torch.from_numpy(z["grad_" + k])
        assert (g[k] - ref).norm() <= 1e-8 * ref.norm() + 1e-12, k


@pytest.mark.parametrize("fixture", ["laikago", "laikago_air", "human", "quad"])
def test_cpu_port_f32_within_north_star_tolerance(fixture):
    """fp32 arithmetic of the shared header vs the float64 oracle: pose <= 1e-4, gradients <= 1e-3 -- except laikago in
    stiff contact, where the bound per gradient is max(1e-3, 2 x the measured fp32 noise floor) (helpers.fp32_noise_floor:
    the autograd oracle itself re-run in float32, an evaluation that shares no code with the port)."""
    from helpers import fp32_noise_floor
    z = np.load(os.path.join(GOLDEN, "rollout_%s.npz" % fixture))
    rm = load_robot(str(z["robot"]))
    d = {k: torch.from_numpy(z["in_" + k]).float() for k in KEYS}
    cpu = CpuRollout(rm)
    pos, vel = cpu.forward(d, float(z["dt"]), int(z["stride"]), int(z["nframes"]))
    assert (pos.double() - torch.from_numpy(z["pos"])).abs().max() < 1e-4
    g = cpu.backward(torch.from_numpy(z["adj_pos"]).float(), torch.from_numpy(z["adj_vel"]).float())
    tol = {k: 1e-3 for k in KEYS}
    if fixture == "laikago":
        floor, _ = fp32_noise_floor(rm, {k: torch.from_numpy(z["in_" + k]) for k in KEYS}, int(z["stride"]),
                                    int(z["nframes"]), adj_pos=torch.from_numpy(z["adj_pos"]),
                                    adj_vel=torch.from_numpy(z["adj_vel"]))
        tol = {k: max(1e-3, 2.0 * floor[k]) for k in KEYS}
    errs = {}
    for k in KEYS:
        ref = torch.from_numpy(z["grad_" + k])
        errs[k] = float((g[k].double() - ref).norm() / ref.norm())
    print("\n[cpu port f32, %s] " % fixture + "; ".join("%s %.1e" % (k, errs[k]) for k in KEYS))
    for k in KEYS:
        assert errs[k] <= tol[k], (k, errs[k], tol[k])


def test_cpu_port_fk_adjoint_matches_autograd():
    from oracle import sim_oracle as so
    from helpers import make_inputs
    for robot in ("laikago", "human"):
        rm, d = make_inputs(robot, bs=5, T=1, seed=8, ang=0.7, qd_std=0.4, quat_noise=0.3, normalize_quat=False)
        q = d["q_init"].clone().requires_grad_(True)
        qd = d["qd_init"].clone().requires_grad_(True)
        bq, bqd = so.eval_fk(so.OracleModel(rm), q, qd)
        g = torch.Generator().manual_seed(1)
        w1, w2 = torch.randn(bq.shape, generator=g, dtype=torch.float64), torch.randn(bqd.shape, generator=g,
                                                                                        dtype=torch.float64)
        gq, gqd = torch.autograd.grad((bq * w1).sum() + (bqd * w2).sum(), [q, qd])
        cpu = CpuRollout(rm)
        cbq, cbqd = cpu.fk(d["q_init"], d["qd_init"])
        assert (cbq - bq.detach()).abs().max() < 1e-12 and (cbqd - bqd.detach()).abs().max() < 1e-12
        aq, aqd = cpu.fk_backward(d["q_init"], d["qd_init"], w1, w2)
        assert (aq - gq).abs().max() < 1e-10 and (aqd - gqd).abs().max() < 1e-10


def test_generic_feature_set_matches_autograd_oracle():
    """FIXED + REVOLUTE + COMPOUND in one tree, active limits, non-identity q_off, spheres / capsules, two materials
    with kd > 0: float64 CPU port (the header the generic CUDA instance compiles) vs the autograd oracle."""
    from oracle import sim_oracle as so
    from helpers import make_inputs, make_mixed_robot, settle_height
    rm = make_mixed_robot()
    assert rm.nqd == 14 and rm.nq == 15 and len(set(rm.contact_material.tolist())) > 1 and rm.contact_dist.max() > 0
    stride, F, bs = 8, 3, 4
    T = stride * (F - 1) + 1
    rm, d = make_inputs(rm, bs=bs, T=T, seed=17, ang=0.25, res_f_std=0.05, torque_std=0.05, lin_vel=0.3, qd_std=0.05)
    d = settle_height(rm, d, 0.004)
    m = so.OracleModel(rm)
    a = {k: d[k].clone().requires_grad_(True) for k in KEYS}
    pos, vel, grf, jaf = so.rollout(m, a["q_init"], a["qd_init"], a["torques"], a["res_f"], a["refs"], a["target_ke"],
                                    a["target_kd"], a["body_inv_mass"], a["body_inertia"], a["body_inv_inertia"], 5e-4,
                                    stride, F)
    assert grf.abs().max() > 1.0  # in contact
    g = torch.Generator().manual_seed(2)
    wp, wv = torch.randn(pos.shape, generator=g, dtype=torch.float64), torch.randn(vel.shape, generator=g,
                                                                                    dtype=torch.float64) * 0.1
    grads = torch.autograd.grad((pos * wp).sum() + (vel * wv).sum(), [a[k] for k in KEYS])
    cpu = CpuRollout(rm)
    p2, v2, g2, j2 = cpu.forward(d, 5e-4, stride, F, want_forces=True)
    assert (p2 - pos.detach()).abs().max() < 1e-11 and (g2 - grf).abs().max() < 1e-6 and (j2 - jaf).abs().max() < 1e-6
    out = cpu.backward(wp, wv)
    for k, gr in zip(KEYS, grads):
        assert (out[k] - gr).norm() <= 1e-9 * gr.norm() + 1e-12, k
