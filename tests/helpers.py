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


def make_mixed_robot():
    """Synthetic articulation exercising everything the three shipped robots do NOT: FIXED joints, REVOLUTE and
    COMPOUND joints in one tree, active joint limits, a non-identity joint_X_c rotation on a COMPOUND joint, sphere /
    capsule contacts (dist > 0), two contact materials with contact damping kd > 0, a non-unit revolute axis and a
    parent with three children. Drives the GENERIC kernel instance (JM_ALL, LIMITS, QOFF)."""
    from ppr_diffphys_b200.model import (ArticulationBuilder, RobotModel, JOINT_FREE, JOINT_REVOLUTE, JOINT_COMPOUND,
                                         JOINT_FIXED, quat_rpy)
    b = ArticulationBuilder()
    m1, m2 = (1e4, 0.0, 1e2, 1.0), (5e3, 50.0, 80.0, 0.6)
    xf = lambda p, rpy: np.concatenate([np.asarray(p, float), quat_rpy(*rpy)])
    r = b.add_body(-1, JOINT_FREE, armature=0.01, name="root")
    b.add_shape_box(r, np.zeros(3), quat_rpy(0, 0, 0), 0.2, 0.05, 0.1, 1000.0, m1)
    a = b.add_body(r, JOINT_REVOLUTE, joint_xform=xf([0.2, -0.05, 0.0], [0.1, 0.2, -0.1]), joint_axis=(0.6, 0.0, 0.8 * 1.5),
                   lower=-0.1, upper=0.1, limit_ke=50.0, limit_kd=1.0, target_ke=100.0, target_kd=2.0, armature=0.01, name="a")
    b.add_shape_box(a, np.array([0.0, -0.1, 0.0]), quat_rpy(0, 0, 0), 0.03, 0.1, 0.03, 1000.0, m1)
    c = b.add_body(a, JOINT_COMPOUND, joint_xform=xf([0.0, -0.2, 0.0], [0.0, 0.1, 0.0]),
                   joint_xform_child=xf([0, 0, 0], [0.06, -0.04, 0.05]), lower=-0.2, upper=0.2, limit_ke=30.0, limit_kd=0.5,
                   target_ke=80.0, target_kd=1.0, armature=0.01, name="c")
    b.add_shape_sphere(c, np.array([0.0, -0.1, 0.0]), quat_rpy(0, 0, 0), 0.05, 1000.0, m2)
    # identity joint rotation: the product's FIXED-joint angle 2*atan2(|e|, w) equals the reference's literal
    # 2*acos(w) only for a UNIT relative quaternion; an f32-rounded joint_X_p quaternion (|q|^2 - 1 ~ 1e-7) makes the
    # literal acos form (not scale invariant, ill-conditioned near identity) drift by (|q|^2 - 1) / angle
    d = b.add_body(r, JOINT_FIXED, joint_xform=xf([-0.2, 0.0, 0.0], [0.0, 0.0, 0.0]), armature=0.01, name="d")
    b.add_shape_box(d, np.zeros(3), quat_rpy(0, 0, 0), 0.05, 0.05, 0.05, 1000.0, m1)
    e = b.add_body(d, JOINT_REVOLUTE, joint_xform=xf([0.0, -0.05, 0.0], [0, 0, 0]), joint_axis=(0.0, 0.0, 1.0),
                   target_ke=100.0, target_kd=2.0, armature=0.01, name="e")
    b.add_shape_capsule(e, np.array([0.0, -0.12, 0.0]), quat_rpy(0, 0, np.pi / 2), 0.03, 0.08, 1000.0, m2)
    f = b.add_body(r, JOINT_COMPOUND, joint_xform=xf([0.0, -0.05, 0.1], [0.2, 0, 0]), target_ke=80.0, target_kd=1.0,
                   armature=0.01, name="f")
    b.add_shape_box(f, np.array([0.0, -0.1, 0.0]), quat_rpy(0, 0, 0), 0.03, 0.1, 0.03, 1000.0, m1)
    b.joint_q[0:7] = [0, 0.4, 0, 0, 0, 0, 1]
    nb = len(b.body_mass)
    for i in range(nb):
        b.body_inertia[i] = b.body_inertia[i] / b.body_mass[i]
    cb, cp, cd, cm = b.collide()
    f32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    i32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.int32))
    ke = [0.0] * 6 + list(b.joint_target_ke[6:])
    kd = [0.0] * 6 + list(b.joint_target_kd[6:])
    return RobotModel(name="mixed", joint_type=i32(b.joint_type), joint_parent=i32(b.joint_parent),
                      joint_X_p=f32(b.joint_X_p), joint_X_c=f32(b.joint_X_c), joint_axis=f32(b.joint_axis),
                      joint_q_start=i32(b.joint_q_start), joint_qd_start=i32(b.joint_qd_start),
                      joint_limit_lower=f32(b.joint_limit_lower), joint_limit_upper=f32(b.joint_limit_upper),
                      joint_limit_ke=f32(b.joint_limit_ke), joint_limit_kd=f32(b.joint_limit_kd),
                      joint_target_ke=f32(ke), joint_target_kd=f32(kd), body_com=f32(b.body_com),
                      body_mass=f32(b.body_mass), norm_body_inertia=f32(b.body_inertia), contact_body=i32(cb),
                      contact_point=f32(cp), contact_dist=f32(cd), contact_material=i32(cm),
                      shape_materials=f32(b.shape_materials), gravity=f32([0.0, -9.80665, 0.0]),
                      joint_q_rest=f32(b.joint_q), joint_attach_ke=4000.0, joint_attach_kd=50.0,
                      body_names=list(b.body_name))


ROLLOUT_KEYS = ["q_init", "qd_init", "torques", "res_f", "refs", "target_ke", "target_kd", "body_inv_mass",
                "body_inertia", "body_inv_inertia"]


def oracle_rollout_grads(rm, d, stride, F, dtype=torch.float64, adj_pos=None, adj_vel=None, loss_fn=None, dt=5e-4,
                         keys=ROLLOUT_KEYS):
    """oracle.sim_oracle rollout in ``dtype`` on the [bs,...] inputs ``d``; gradients of <adj_pos, pos> + <adj_vel, vel>
    (or of ``loss_fn(pos, vel)``) w.r.t. ``keys``.  Returns pos, vel, {key: grad} as float64 CPU tensors."""
    from oracle import sim_oracle as so
    m = so.OracleModel(rm, dtype=dtype)
    a = {k: d[k].to(dtype).clone().requires_grad_(k in keys) for k in ROLLOUT_KEYS}
    pos, vel = so.rollout(m, a["q_init"], a["qd_init"], a["torques"], a["res_f"], a["refs"], a["target_ke"], a["target_kd"],
                          a["body_inv_mass"], a["body_inertia"], a["body_inv_inertia"], dt, stride, F)[:2]
    if loss_fn is not None:
        loss = loss_fn(pos, vel)
    else:
        loss = (pos * adj_pos.to(dtype).reshape(pos.shape)).sum() + (vel * adj_vel.to(dtype).reshape(vel.shape)).sum()
    grads = torch.autograd.grad(loss, [a[k] for k in keys], allow_unused=True)
    out = {k: (torch.zeros_like(a[k]) if g is None else g).detach().double() for k, g in zip(keys, grads)}
    return pos.detach().double(), vel.detach().double(), out


def rel_err(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def fp32_noise_floor(rm, d, stride, F, **kw):
    """Per-key relative difference between an independent float32 evaluation (the autograd oracle run in float32) and
    the float64 oracle on the same inputs: what single precision can resolve for this problem, whatever the code.
    Returns ({key: floor}, float64 grads)."""
    _, _, g64 = oracle_rollout_grads(rm, d, stride, F, dtype=torch.float64, **kw)
    _, _, g32 = oracle_rollout_grads(rm, d, stride, F, dtype=torch.float32, **kw)
    return {k: rel_err(g32[k], g64[k]) for k in g64}, g64
