"""Shared synthetic-input builders for oracle / CUDA parity tests (seeded, deterministic)."""
import numpy as np
import torch

from ppr_diffphys_b200 import load_robot


def make_inputs(robot, bs, T, seed=0, height=None, ang=0.2, qd_std=0.1, ref_amp=0.1, dtype=torch.float64,
                res_f_std=0.0, torque_std=0.0, quat_noise=0.01, normalize_quat=True, lin_vel=0.0):
    """Synthetic rollout inputs in the [bs,...] layout of oracle.sim_oracle.rollout
    (SURVEY.md section 8d configs 3-5 distributions)."""
    rm = load_robot(robot) if isinstance(robot, str) else robot
    g = torch.Generator().manual_seed(seed)
    nb, nq, nqd = rm.nb, rm.nq, rm.nqd
    B = nqd - 6
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    ru = lambda *s: torch.rand(*s, generator=g, dtype=torch.float64)
    ja = (ru(bs, B) * 2 - 1) * ang
    quat = torch.tensor([0.0, 0, 0, 1.0]).expand(bs, 4) + rn(bs, 4) * quat_noise
    if normalize_quat:
        quat = quat / quat.norm(dim=-1, keepdim=True)
    pos = torch.zeros(bs, 3, dtype=torch.float64)
    pos[:, 1] = 0.45 if height is None else height
    q_init = torch.cat([pos, quat, ja], -1)
    qd_init = rn(bs, nqd) * qd_std
    if lin_vel > 0:
        qd_init[:, 3] = (ru(bs) * 2 - 1) * lin_vel
        qd_init[:, 5] = (ru(bs) * 2 - 1) * lin_vel
    t = torch.arange(T, dtype=torch.float64)[:, None, None]
    phase = ru(1, bs, B) * 2 * np.pi
    refs = torch.zeros(T, bs, nqd, dtype=torch.float64)
    refs[:, :, 6:] = ja[None] + ref_amp * torch.sin(2 * np.pi * t / 64.0 + phase)
    torques = torch.zeros(T, bs, nqd, dtype=torch.float64)
    torques[:, :, 6:] = rn(T, bs, B) * torque_std
    res_f = rn(T, bs, nb, 6) * res_f_std
    ke = torch.as_tensor(rm.joint_target_ke, dtype=torch.float64)[None].repeat(bs, 1)
    kd = torch.as_tensor(rm.joint_target_kd, dtype=torch.float64)[None].repeat(bs, 1)
    mass = torch.as_tensor(rm.body_mass, dtype=torch.float64)[None].repeat(bs, 1)
    mass = mass * (1.0 + 0.1 * (ru(bs, nb) - 0.5))
    nI = torch.as_tensor(rm.norm_body_inertia, dtype=torch.float64)[None]
    inv_m = 1.0 / mass
    I = nI * mass[..., None, None]
    inv_I = torch.linalg.inv(I)
    d = dict(q_init=q_init, qd_init=qd_init, torques=torques, res_f=res_f, refs=refs, target_ke=ke, target_kd=kd,
             body_mass=mass, body_inv_mass=inv_m, body_inertia=I, body_inv_inertia=inv_I)
    return rm, {k: v.to(dtype).contiguous() for k, v in d.items()}


def standing_height(rm, q_rot=None, margin=1e-3, ja=None):
    """Root height such that the lowest contact point sits ``margin`` above the ground at zero pose."""
    from oracle.sim_oracle import OracleModel, eval_fk, transform_point
    m = OracleModel(rm)
    q = torch.zeros(1, rm.nq, dtype=torch.float64)
    q[0, 6] = 1.0
    if ja is not None:
        q[0, 7:] = ja
    bq, _ = eval_fk(m, q, torch.zeros(1, rm.nqd, dtype=torch.float64))
    cp = transform_point(bq[:, m.contact_body], m.contact_point[None])
    return float(-(cp[..., 1] - m.contact_dist[None]).min() + margin)


def settle_height(rm, d, penetration=0.003):
    """Shift every env's root height so that its lowest contact point penetrates the ground by ``penetration``
    (uses the actual initial joint angles / root orientation)."""
    from oracle.sim_oracle import OracleModel, eval_fk, transform_point
    m = OracleModel(rm)
    q = d["q_init"].double().clone()
    bq, _ = eval_fk(m, q, torch.zeros(q.shape[0], rm.nqd, dtype=torch.float64))
    cp = transform_point(bq[:, m.contact_body], m.contact_point[None])
    low = (cp[..., 1] - m.contact_dist[None]).min(dim=1)[0]
    q[:, 1] = q[:, 1] - low - penetration
    d["q_init"] = q.to(d["q_init"].dtype)
    return d
