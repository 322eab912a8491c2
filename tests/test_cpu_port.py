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


@pytest.mark.parametrize("fixture", ["laikago_air", "human", "quad"])
def test_cpu_port_f32_within_north_star_tolerance(fixture):
    """fp32 arithmetic of the shared header vs the float64 oracle: pose <= 1e-4, gradients <= 1e-3."""
    z = np.load(os.path.join(GOLDEN, "rollout_%s.npz" % fixture))
    rm = load_robot(str(z["robot"]))
    d = {k: torch.from_numpy(z["in_" + k]).float() for k in KEYS}
    cpu = CpuRollout(rm)
    pos, vel = cpu.forward(d, float(z["dt"]), int(z["stride"]), int(z["nframes"]))
    assert (pos.double() - torch.from_numpy(z["pos"])).abs().max() < 1e-4
    g = cpu.backward(torch.from_numpy(z["adj_pos"]).float(), torch.from_numpy(z["adj_vel"]).float())
    for k in KEYS:
        ref = torch.from_numpy(z["grad_" + k])
        assert (g[k].double() - ref).norm() <= 1e-3 * ref.norm(), k


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
