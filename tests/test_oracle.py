"""Pins the float64 oracle without the (uninstallable) Warp reference: finite differences,
analytic free fall, momentum conservation, FK <-> joint-error round trip."""
import numpy as np
import pytest
import torch

from oracle import sim_oracle as so
from ppr_diffphys_b200 import load_robot
from helpers import make_inputs, standing_height

ROBOTS = ["laikago", "human", "quad"]
DT = 5e-4


@pytest.mark.parametrize("robot", ROBOTS)
def test_fk_joint_roundtrip(robot):
    """A FK-consistent state with refs == joint angles and zero velocity produces zero joint wrench:
    pins REVOLUTE twist extraction and the COMPOUND x-y'-z'' convention (integrator_euler.py:413-429)."""
    rm, d = make_inputs(robot, bs=4, T=1, seed=1, ang=0.9, qd_std=0.0, height=2.0)
    m = so.OracleModel(rm)
    bq, bqd = so.eval_fk(m, d["q_init"], d["qd_init"])
    refs = torch.zeros(4, rm.nqd, dtype=torch.float64)
    refs[:, 6:] = d["q_init"][:, 7:]
    w = so.eval_body_joints(m, bq, bqd, refs, torch.zeros_like(refs), d["target_ke"], d["target_kd"])
    assert w.abs().max() < 2e-3  # f32-rounded joint_X_p quaternions are unit only to 1e-7; ke ~ 660
    # and a perturbed target produces exactly ke * delta about the joint axis
    refs2 = refs.clone()
    refs2[:, 6:] += 0.1
    w2 = so.eval_body_joints(m, bq, bqd, refs2, torch.zeros_like(refs), d["target_ke"], d["target_kd"])
    assert w2.abs().max() > 1.0


@pytest.mark.parametrize("robot", ROBOTS)
def test_fk_velocity_is_time_derivative(robot):
    """body_qd angular part equals the finite-difference angular velocity of body_q (pins the FK twist rule
    up to the documented ``w x com_local`` linear term)."""
    rm, d = make_inputs(robot, bs=2, T=1, seed=2, ang=0.5, qd_std=0.3, height=2.0)
    m = so.OracleModel(rm)
    q, qd = d["q_init"], d["qd_init"].clone()
    qd[:, :6] = 0  # FREE-joint coordinates are not integrable by simple addition
    h = 1e-6
    qp, qm = q.clone(), q.clone()
    qp[:, 7:] += h * qd[:, 6:]
    qm[:, 7:] -= h * qd[:, 6:]
    b0, bd = so.eval_fk(m, q, qd)
    bp, _ = so.eval_fk(m, qp, qd)
    bm, _ = so.eval_fk(m, qm, qd)
    dq = (bp[..., 3:7] - bm[..., 3:7]) / (2 * h)
    # w = 2 * dq * conj(q)
    w_fd = 2.0 * so.quat_mul(dq, so.quat_inverse(b0[..., 3:7]))[..., :3]
    assert torch.allclose(w_fd, bd[..., :3], atol=1e-6)


def test_free_fall_analytic():
    rm, d = make_inputs("laikago", bs=1, T=40, seed=0, ang=0.0, qd_std=0.0, height=3.0, ref_amp=0.0, quat_noise=0.0)
    m = so.OracleModel(rm)
    F, stride = 2, 32
    pos, vel, _, _ = so.rollout(m, d["q_init"], d["qd_init"], d["torques"][:33], d["res_f"][:33], d["refs"][:33],
                                d["target_ke"], d["target_kd"], d["body_inv_mass"], d["body_inertia"],
                                d["body_inv_inertia"], DT, stride, F)
    g = float(rm.gravity[1])
    n = 32
    assert torch.allclose(vel[1, 0, :, 4], torch.full((rm.nb,), g * n * DT, dtype=torch.float64), atol=1e-9)
    dy = pos[1, 0, :, 1] - pos[0, 0, :, 1]
    assert torch.allclose(dy, torch.full((rm.nb,), g * DT * DT * n * (n + 1) / 2, dtype=torch.float64), atol=1e-9)


@pytest.mark.parametrize("robot", ROBOTS)
def test_linear_momentum_conserved_without_gravity_and_contact(robot):
    rm, d = make_inputs(robot, bs=2, T=33, seed=3, ang=0.4, qd_std=0.2, height=5.0)
    rm.gravity = np.zeros(3, dtype=np.float32)
    m = so.OracleModel(rm)
    pos, vel, _, _ = so.rollout(m, d["q_init"], d["qd_init"], d["torques"], d["res_f"], d["refs"], d["target_ke"],
                                d["target_kd"], d["body_inv_mass"], d["body_inertia"], d["body_inv_inertia"], DT, 32, 2)
    p = (vel[..., 3:] * d["body_mass"][None, ..., None]).sum(2)
    assert torch.allclose(p[0], p[1], atol=1e-9)
    assert (vel[1] - vel[0]).abs().max() > 1e-3  # something actually happened


def test_contact_normal_and_friction_cap():
    rm = load_robot("human")
    m = so.OracleModel(rm)
    h = standing_height(rm, margin=-0.004)  # lowest corners 4 mm under ground
    q = torch.zeros(1, rm.nq, dtype=torch.float64)
    q[0, 1], q[0, 6] = h, 1.0
    qd = torch.zeros(1, rm.nqd, dtype=torch.float64)
    qd[0, 3] = 1.0  # slide along +x
    bq, bqd = so.eval_fk(m, q, qd)
    w = so.eval_body_contacts(m, bq, bqd)  # wrench subtracted from body_f
    cp = so.transform_point(bq[:, m.contact_body], m.contact_point[None])[0]
    pen = cp[:, 1].clamp(max=0.0)
    ke, kf, mu = 1e4, 1e2, 1.0
    fy = (w[0, :, 4]).sum()
    assert torch.allclose(fy, (ke * pen).sum(), rtol=1e-9)            # fn = c*ke (negative); body_f -= f
    fx = w[0, :, 3].sum()
    cap = torch.minimum(torch.full_like(pen, kf * 1.0), -mu * ke * pen)[pen < 0].sum()
    assert torch.allclose(fx, cap, rtol=1e-9)                          # friction along +vt, later subtracted


def _loss_fn(m, d, keys, stride, F, wpos, wvel):
    def f(*args):
        dd = dict(d)
        for k, a in zip(keys, args):
            dd[k] = a
        if "body_mass" in keys:  # chain mass -> inv_mass, I, inv_I like dp_model.py:725-730
            nI = d["_nI"]
            dd["body_inv_mass"] = 1.0 / dd["body_mass"]
            dd["body_inertia"] = nI * dd["body_mass"][..., None, None]
            dd["body_inv_inertia"] = torch.linalg.inv(dd["body_inertia"])
        pos, vel = so.rollout(m, dd["q_init"], dd["qd_init"], dd["torques"], dd["res_f"], dd["refs"], dd["target_ke"],
                              dd["target_kd"], dd["body_inv_mass"], dd["body_inertia"], dd["body_inv_inertia"], DT,
                              stride, F)[:2]
        return (pos * wpos).sum() + (vel * wvel).sum()
    return f


@pytest.mark.parametrize("robot,height", [("laikago", None), ("human", "contact"), ("quad", "contact")])
def test_autograd_matches_finite_differences(robot, height):
    stride, F = 8, 3
    T = stride * (F - 1) + 1
    rm = load_robot(robot)
    hgt = standing_height(rm, margin=-0.003) if height == "contact" else 0.45
    rm, d = make_inputs(rm, bs=2, T=T, seed=4, height=hgt, res_f_std=0.1, torque_std=0.1, lin_vel=0.5)
    m = so.OracleModel(rm)
    d["_nI"] = torch.as_tensor(rm.norm_body_inertia, dtype=torch.float64)[None]
    g = torch.Generator().manual_seed(5)
    wpos = torch.randn(F, 2, rm.nb, 7, generator=g, dtype=torch.float64)
    wvel = torch.randn(F, 2, rm.nb, 6, generator=g, dtype=torch.float64) * 0.1
    keys = ["q_init", "qd_init", "refs", "torques", "res_f", "target_ke", "target_kd", "body_mass"]
    f = _loss_fn(m, d, keys, stride, F, wpos, wvel)
    args = [d[k].clone().requires_grad_(True) for k in keys]
    loss = f(*args)
    grads = torch.autograd.grad(loss, args)
    for k, a, gr in zip(keys, args, grads):
        assert torch.isfinite(gr).all(), k
        for trial in range(2):
            v = torch.randn(a.shape, generator=g, dtype=torch.float64)
            v = v / v.norm()
            h = 1e-6 * max(1.0, float(a.abs().max()))
            ap = [x.detach() for x in args]
            am = [x.detach() for x in args]
            i = keys.index(k)
            ap[i] = ap[i] + h * v
            am[i] = am[i] - h * v
            fd = (f(*ap) - f(*am)) / (2 * h)
            an = (gr * v).sum()
            assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)) + 1e-7, (k, float(fd), float(an))


@pytest.mark.parametrize("fixture", ["laikago", "laikago_air", "human", "quad"])
def test_oracle_reproduces_committed_golden_vectors(fixture):
    """The fixtures under tests/golden/ (what the GPU parity tests compare against) are exactly what the oracle
    computes today: guards against silent drift of the oracle after the fixtures were frozen."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_%s.npz" % fixture))
    keys = ["q_init", "qd_init", "torques", "res_f", "refs", "target_ke", "target_kd", "body_inv_mass", "body_inertia",
            "body_inv_inertia"]
    m = so.OracleModel(load_robot(str(z["robot"])))
    o = {k: torch.from_numpy(z["in_" + k]).double().requires_grad_(True) for k in keys}
    pos, vel, grf, jaf = so.rollout(m, *[o[k] for k in keys], float(z["dt"]), int(z["stride"]), int(z["nframes"]))
    assert (pos.detach() - torch.from_numpy(z["pos"])).abs().max() < 1e-12
    assert (vel.detach() - torch.from_numpy(z["vel"])).abs().max() < 1e-11
    assert (grf.detach() - torch.from_numpy(z["grf"])).abs().max() < 1e-9
    assert (jaf.detach() - torch.from_numpy(z["jaf"])).abs().max() < 1e-9
    loss = (pos * torch.from_numpy(z["adj_pos"])).sum() + (vel * torch.from_numpy(z["adj_vel"])).sum()
    grads = torch.autograd.grad(loss, [o[k] for k in keys])
    for k, g in zip(keys, grads):
        ref = torch.from_numpy(z["grad_" + k])
        assert (g - ref).norm() <= 1e-10 * ref.norm() + 1e-14, k


def test_fp32_noise_floor_of_the_reference_formulation():
    """What single precision resolves on the committed fixtures, measured with the oracle itself (float32 vs float64 run
    of the same autograd restatement): the GPU parity tests bound the CUDA gradients of laikago-in-contact by
    max(1e-3, 2 x this floor) instead of a flat loose tolerance.  Pinned here so that a change of the oracle that moved
    the floor (and with it the GPU tolerance) is seen on the CPU."""
    import os

    import numpy as np
    from helpers import ROLLOUT_KEYS, fp32_noise_floor
    from ppr_diffphys_b200 import load_robot
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for fixture, lo, hi in (("laikago", 3e-4, 1e-2), ("human", 0.0, 1e-4)):
        z = np.load(os.path.join(golden, "rollout_%s.npz" % fixture))
        rm = load_robot(str(z["robot"]))
        d = {k: torch.from_numpy(z["in_" + k]) for k in ROLLOUT_KEYS}
        floor, _ = fp32_noise_floor(rm, d, int(z["stride"]), int(z["nframes"]), adj_pos=torch.from_numpy(z["adj_pos"]),
                                    adj_vel=torch.from_numpy(z["adj_vel"]))
        print("\n[fp32 floor, %s] " % fixture + "; ".join("%s %.1e" % kv for kv in floor.items()))
        assert all(lo <= v <= hi for v in floor.values()), (fixture, floor)
