"""ORACLE (test infrastructure, not product code) -- PARITY UNPINNED.

CPU restatement, in vectorised float64 PyTorch with autograd supplying the adjoint, of the
reference's motion-imitation rollout hot path:

  eval_fk                 third-party warp_lang==0.7.2 ``warp.sim.articulation.eval_articulation_fk``
                          (NOT under /root/reference; call sites diffphys/dp_model.py:1068,1204) --
                          restated from its published source, UNVERIFIED here
  wp_add                  diffphys/dp_model.py:1133-1142
  eval_body_contacts      diffphys/integrator_euler.py:93-179
  quat_twist / quat_decompose / eval_joint_force   diffphys/integrator_euler.py:234-286
  eval_body_joints        diffphys/integrator_euler.py:289-451
  integrate_bodies        diffphys/integrator_euler.py:21-91
  compute_forces / simulate (launch order, grf / joint_f side channels)   :491-551, :579-620
  ForwardWarp rollout schedule + which gradients exist   diffphys/dp_model.py:1145-1400

"Parity unpinned": the reference ships no tests / golden vectors, and Warp cannot be installed in
this environment (no network), so this restatement cannot be checked against reference outputs.
It is pinned instead by (tests/test_oracle.py): central finite differences in float64, analytic
free fall, static contact force balance, momentum conservation, and the FK o joint-error-extraction
round trip, which fixes the COMPOUND convention because integrator_euler.py:413-429 is the in-tree
inverse of the out-of-tree FK.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Conventions: quaternions xyzw, transform (p, q), spatial vector
(angular, linear), ground plane y = 0 with normal +Y.
"""
from __future__ import annotations

import torch

JOINT_PRISMATIC, JOINT_REVOLUTE, JOINT_BALL, JOINT_FIXED, JOINT_FREE, JOINT_COMPOUND, JOINT_UNIVERSAL = range(7)


# --------------------------------------------------------------------------- Warp built-ins
def quat_mul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + bw * ax + ay * bz - az * by,
                        aw * by + bw * ay + az * bx - ax * bz,
                        aw * bz + bw * az + ax * by - ay * bx,
                        aw * bw - ax * bx - ay * by - az * bz], -1)


def quat_inverse(q):  # Warp: conjugate
    return torch.cat([-q[..., :3], q[..., 3:]], -1)


def quat_rotate(q, v):
    """Warp ``quat_rotate`` -- NOT normalising: v(2w^2-1) + 2w(u x v) + 2u(u.v)."""
    u, w = q[..., :3], q[..., 3:]
    return v * (2.0 * w * w - 1.0) + 2.0 * w * torch.cross(u, v, dim=-1) + 2.0 * u * (u * v).sum(-1, keepdim=True)


def quat_rotate_inv(q, v):
    u, w = q[..., :3], q[..., 3:]
    return v * (2.0 * w * w - 1.0) - 2.0 * w * torch.cross(u, v, dim=-1) + 2.0 * u * (u * v).sum(-1, keepdim=True)


def quat_from_axis_angle(axis, angle):
    h = 0.5 * angle
    return torch.cat([axis * torch.sin(h)[..., None], torch.cos(h)[..., None]], -1)


def transform_point(X, p):
    return X[..., :3] + quat_rotate(X[..., 3:7], p)


def transform_mul(A, B):
    return torch.cat([A[..., :3] + quat_rotate(A[..., 3:7], B[..., :3]), quat_mul(A[..., 3:7], B[..., 3:7])], -1)


def safe_length(v):
    l2 = (v * v).sum(-1)
    ok = l2 > 0
    return torch.where(ok, torch.sqrt(torch.where(ok, l2, torch.ones_like(l2))), torch.zeros_like(l2))


def safe_normalize(v):
    """Warp ``normalize``: v/|v|, zero vector (and zero adjoint) at |v| = 0."""
    l2 = (v * v).sum(-1, keepdim=True)
    ok = l2 > 0
    inv = torch.where(ok, torch.rsqrt(torch.where(ok, l2, torch.ones_like(l2))), torch.zeros_like(l2))
    return v * inv


def safe_acos(x):
    """acos clamped to [-1,1]; adjoint 0 at saturation (the reference scrubs the NaN, dp_utils.py:53)."""
    ok = x.abs() < 1.0
    xs = torch.where(ok, x, torch.zeros_like(x))
    sat = torch.where(x > 0, torch.zeros_like(x), torch.full_like(x, torch.pi))
    return torch.where(ok, torch.acos(xs), sat)


def safe_asin(x):
    ok = x.abs() < 1.0
    xs = torch.where(ok, x, torch.zeros_like(x))
    sat = torch.where(x > 0, torch.full_like(x, 0.5 * torch.pi), torch.full_like(x, -0.5 * torch.pi))
    return torch.where(ok, torch.asin(xs), sat)


def wp_min(a, b):  # adjoint goes to a iff a < b
    return torch.where(a < b, a, b)


def wp_clamp(x, lo, hi):  # adjoint passes iff lo <= x <= hi
    return torch.where(x < lo, torch.full_like(x, lo), torch.where(x > hi, torch.full_like(x, hi), x))


def wp_sign(x):
    return torch.where(x < 0, -torch.ones_like(x), torch.ones_like(x))


# --------------------------------------------------------------------------- static model
class OracleModel:
    """Torch view of a ppr_diffphys_b200.model.RobotModel (any object with those attributes)."""

    def __init__(self, rm, dtype=torch.float64, device="cpu"):
        t = lambda a: torch.as_tensor(a, dtype=dtype, device=device)
        self.dtype, self.device = dtype, device
        self.nb, self.nq, self.nqd, self.nc = rm.nb, rm.nq, rm.nqd, rm.nc
        self.joint_type = [int(x) for x in rm.joint_type]
        self.joint_parent = [int(x) for x in rm.joint_parent]
        self.joint_q_start = [int(x) for x in rm.joint_q_start]
        self.joint_qd_start = [int(x) for x in rm.joint_qd_start]
        self.joint_X_p = t(rm.joint_X_p)
        self.joint_X_c = t(rm.joint_X_c)
        self.joint_axis = t(rm.joint_axis)
        self.joint_limit_lower = t(rm.joint_limit_lower)
        self.joint_limit_upper = t(rm.joint_limit_upper)
        self.joint_limit_ke = t(rm.joint_limit_ke)
        self.joint_limit_kd = t(rm.joint_limit_kd)
        self.body_com = t(rm.body_com)
        self.contact_body = torch.as_tensor(rm.contact_body, dtype=torch.long, device=device)
        self.contact_point = t(rm.contact_point)
        self.contact_dist = t(rm.contact_dist)
        self.contact_mat = t(rm.shape_materials)[torch.as_tensor(rm.contact_material, dtype=torch.long)]  # [nc,4]
        self.gravity = t(rm.gravity)
        self.joint_attach_ke = float(rm.joint_attach_ke)
        self.joint_attach_kd = float(rm.joint_attach_kd)
        self.parent_idx = torch.as_tensor([max(p, 0) for p in self.joint_parent], dtype=torch.long, device=device)
        self.has_parent = torch.as_tensor([p >= 0 for p in self.joint_parent], device=device)


# --------------------------------------------------------------------------- K1: eval_fk
def eval_fk(m: OracleModel, joint_q, joint_qd):
    """joint_q [bs,nq], joint_qd [bs,nqd] -> body_q [bs,nb,7], body_qd [bs,nb,6].
    Restates warp 0.7.2 ``eval_articulation_fk`` (third-party; see module docstring)."""
    bs = joint_q.shape[0]
    z3 = torch.zeros(bs, 3, dtype=joint_q.dtype, device=joint_q.device)
    ident = torch.cat([z3, z3[:, :1], z3[:, :1], z3[:, :1], torch.ones_like(z3[:, :1])], -1)
    body_q, body_qd = [None] * m.nb, [None] * m.nb
    ex = torch.tensor([1.0, 0.0, 0.0], dtype=joint_q.dtype, device=joint_q.device).expand(bs, 3)
    ey = torch.tensor([0.0, 1.0, 0.0], dtype=joint_q.dtype, device=joint_q.device).expand(bs, 3)
    ez = torch.tensor([0.0, 0.0, 1.0], dtype=joint_q.dtype, device=joint_q.device).expand(bs, 3)
    for i in range(m.nb):
        p = m.joint_parent[i]
        X_wp = body_q[p] if p >= 0 else ident
        v_wp = body_qd[p] if p >= 0 else torch.zeros(bs, 6, dtype=joint_q.dtype, device=joint_q.device)
        X_pj = m.joint_X_p[:, i] if m.joint_X_p.dim() == 3 else m.joint_X_p[i].expand(bs, 7)   # per-env [bs,nb,7] or shared
        qs, qds, jt = m.joint_q_start[i], m.joint_qd_start[i], m.joint_type[i]
        if jt == JOINT_FREE:
            X_jc = joint_q[:, qs:qs + 7]
            w_j, v_j = joint_qd[:, qds:qds + 3], joint_qd[:, qds + 3:qds + 6]
        elif jt == JOINT_REVOLUTE:
            axis = m.joint_axis[i].expand(bs, 3)
            X_jc = torch.cat([z3, quat_from_axis_angle(axis, joint_q[:, qs])], -1)
            w_j, v_j = axis * joint_qd[:, qds:qds + 1], z3
        elif jt == JOINT_COMPOUND:
            q0 = quat_from_axis_angle(ex, joint_q[:, qs + 0])
            a1 = quat_rotate(q0, ey)
            q1 = quat_from_axis_angle(a1, joint_q[:, qs + 1])
            a2 = quat_rotate(quat_mul(q1, q0), ez)
            q2 = quat_from_axis_angle(a2, joint_q[:, qs + 2])
            X_jc = torch.cat([z3, quat_mul(q2, quat_mul(q1, q0))], -1)
            w_j = ex * joint_qd[:, qds:qds + 1] + a1 * joint_qd[:, qds + 1:qds + 2] + a2 * joint_qd[:, qds + 2:qds + 3]
            v_j = z3
        elif jt == JOINT_FIXED:
            X_jc, w_j, v_j = ident, z3, z3
        else:
            raise NotImplementedError("joint type %d" % jt)
        X_wj = transform_mul(X_wp, X_pj)
        X_wc = transform_mul(X_wj, X_jc)
        w_w = quat_rotate(X_wj[:, 3:7], w_j)
        v_w = quat_rotate(X_wj[:, 3:7], v_j)
        com = m.body_com[i].expand(bs, 3)
        body_q[i] = X_wc
        body_qd[i] = v_wp + torch.cat([w_w, v_w + torch.cross(w_w, com, dim=-1)], -1)
    return torch.stack(body_q, 1), torch.stack(body_qd, 1)


# --------------------------------------------------------------------------- K3: contacts
def eval_body_contacts(m: OracleModel, body_q, body_qd):
    """Returns the wrench [bs,nb,6] that K3 SUBTRACTS from body_f (integrator_euler.py:179)."""
    bs = body_q.shape[0]
    cb = m.contact_body
    X = body_q[:, cb]            # [bs,nc,7]
    tw = body_qd[:, cb]
    w, v = tw[..., :3], tw[..., 3:]
    n = torch.tensor([0.0, 1.0, 0.0], dtype=body_q.dtype, device=body_q.device)
    cp = transform_point(X, m.contact_point[None]) - n * m.contact_dist[None, :, None]
    r = cp - transform_point(X, m.body_com[cb][None])
    dpdt = v + torch.cross(w, r, dim=-1)
    c = cp[..., 1]
    active = ~(c > 0.0)
    ke, kd, kf, mu = m.contact_mat.unbind(-1)
    vn = dpdt[..., 1]
    vt = dpdt - n * vn[..., None]
    fn = c * ke
    step_c = (c < 0.0).to(c.dtype)
    fd = wp_min(vn, torch.zeros_like(vn)) * kd * step_c
    ft = safe_normalize(vt) * wp_min(kf * safe_length(vt), 0.0 - mu * (fn + fd))[..., None]
    f_total = n * (fn + fd)[..., None] + ft
    f_total = wp_clamp(f_total, -500.0, 500.0)
    f_total = torch.where(active[..., None], f_total, torch.zeros_like(f_total))
    t_total = torch.cross(r, f_total, dim=-1)
    wrench = torch.cat([t_total, f_total], -1)  # [bs,nc,6]
    out = torch.zeros(bs, m.nb, 6, dtype=body_q.dtype, device=body_q.device)
    return out.index_add(1, cb, wrench)


# --------------------------------------------------------------------------- K4: joints
def quat_twist(axis, q):
    a = (q[..., :3] * axis).sum(-1, keepdim=True) * axis
    return safe_normalize(torch.cat([a, q[..., 3:]], -1))


def quat_decompose(q):
    e = torch.eye(3, dtype=q.dtype, device=q.device)
    v0 = quat_rotate(q, e[0].expand(q.shape[:-1] + (3,)))
    v1 = quat_rotate(q, e[1].expand(q.shape[:-1] + (3,)))
    v2 = quat_rotate(q, e[2].expand(q.shape[:-1] + (3,)))
    # wp.mat33(v0,v1,v2) takes COLUMNS: R[i,j] = v_j[i]
    phi = torch.atan2(v2[..., 1], v2[..., 2])
    theta = safe_asin(-v2[..., 0])
    psi = torch.atan2(v1[..., 0], v0[..., 0])
    return -torch.stack([phi, theta, psi], -1)


def eval_joint_force(q, qd, target, ke, kd, act, lo, hi, lke, lkd):
    """Scalar part of eval_joint_force (integrator_euler.py:262-286); caller multiplies by the axis."""
    zero = torch.zeros_like(q)
    lim = torch.where(q < lo, lke * (lo - q) - lkd * wp_min(qd, zero), zero)
    lim = torch.where(q > hi, lke * (hi - q) - lkd * torch.where(qd > zero, qd, zero), lim)
    return ke * (q - target) + kd * qd + act - lim


def eval_body_joints(m: OracleModel, body_q, body_qd, joint_target, joint_act, target_ke, target_kd):
    """Returns the wrench [bs,nb,6] K4 ADDS to body_f (parent +, child -).
    joint_target / joint_act / target_ke / target_kd: [bs,nqd]."""
    bs, nb = body_q.shape[0], m.nb
    dt_, dev = body_q.dtype, body_q.device
    hp = m.has_parent[None, :, None].to(dt_)
    Xp_body = body_q[:, m.parent_idx]                                  # [bs,nb,7]
    X_pj = m.joint_X_p if m.joint_X_p.dim() == 3 else m.joint_X_p[None].expand(bs, nb, 7)
    X_wp = torch.where(m.has_parent[None, :, None], transform_mul(Xp_body, X_pj), X_pj)
    r_p = (X_wp[..., :3] - transform_point(Xp_body, m.body_com[m.parent_idx][None])) * hp
    tw_p = body_qd[:, m.parent_idx] * hp
    w_p, v_p = tw_p[..., :3], tw_p[..., 3:]
    X_wc = body_q
    r_c = X_wc[..., :3] - transform_point(body_q, m.body_com[None])
    w_c, v_c = body_qd[..., :3], body_qd[..., 3:]
    x_p, x_c, q_p, q_c = X_wp[..., :3], X_wc[..., :3], X_wp[..., 3:7], X_wc[..., 3:7]
    x_err = x_c - x_p
    r_err = quat_mul(quat_inverse(q_p), q_c)
    v_err = v_c - v_p
    w_err = w_c - w_p
    ake, akd = m.joint_attach_ke, m.joint_attach_kd
    ads = 0.01
    t_list, f_list = [], []
    for j in range(nb):
        jt, s = m.joint_type[j], m.joint_qd_start[j]
        xe, re, ve, we = x_err[:, j], r_err[:, j], v_err[:, j], w_err[:, j]
        z3 = torch.zeros(bs, 3, dtype=dt_, device=dev)
        if jt == JOINT_FREE:
            t_tot, f_tot = z3, z3
        elif jt == JOINT_FIXED:
            ang = safe_normalize(re[:, :3]) * (safe_acos(re[:, 3]) * 2.0)[:, None]
            f_tot = xe * ake + ve * akd
            t_tot = quat_rotate(q_p[:, j], ang) * ake + we * akd * ads
        elif jt == JOINT_REVOLUTE:
            axis = m.joint_axis[j].expand(bs, 3)
            axis_p = quat_rotate(q_p[:, j], axis)
            axis_c = quat_rotate(q_c[:, j], axis)
            tw = quat_twist(axis, re)
            q = safe_acos(tw[:, 3]) * 2.0 * wp_sign((axis * tw[:, :3]).sum(-1))
            qd = (we * axis_p).sum(-1)
            sc = eval_joint_force(q, qd, joint_target[:, s], target_ke[:, s], target_kd[:, s], joint_act[:, s],
                                  m.joint_limit_lower[s], m.joint_limit_upper[s], m.joint_limit_ke[s],
                                  m.joint_limit_kd[s])
            t_tot = sc[:, None] * axis_p
            swing = torch.cross(axis_p, axis_c, dim=-1)
            f_tot = xe * ake + ve * akd
            t_tot = t_tot + swing * ake + (we - qd[:, None] * axis_p) * akd * ads
        elif jt == JOINT_COMPOUND:
            q_off = m.joint_X_c[j, 3:7].expand(bs, 4)
            q_pc = quat_mul(quat_mul(quat_mul(quat_inverse(q_off), quat_inverse(q_p[:, j])), q_c[:, j]), q_off)
            ang = quat_decompose(q_pc)
            e = torch.eye(3, dtype=dt_, device=dev)
            a0 = e[0].expand(bs, 3)
            q0 = quat_from_axis_angle(a0, ang[:, 0])
            a1 = quat_rotate(q0, e[1].expand(bs, 3))
            q1 = quat_from_axis_angle(a1, ang[:, 1])
            a2 = quat_rotate(quat_mul(q1, q0), e[2].expand(bs, 3))
            q_w = quat_mul(q_p[:, j], q_off)
            t_tot = z3
            for k, ak in enumerate((a0, a1, a2)):
                aw = quat_rotate(q_w, ak)
                sc = eval_joint_force(ang[:, k], (aw * we).sum(-1), joint_target[:, s + k], target_ke[:, s + k],
                                      target_kd[:, s + k], joint_act[:, s + k], m.joint_limit_lower[s + k],
                                      m.joint_limit_upper[s + k], m.joint_limit_ke[s + k], m.joint_limit_kd[s + k])
                t_tot = t_tot + sc[:, None] * aw
            t_tot = wp_clamp(t_tot, -1e4, 1e4)
            f_tot = wp_clamp(xe * ake + ve * akd, -1e4, 1e4)
        else:
            raise NotImplementedError("joint type %d" % jt)
        t_list.append(t_tot)
        f_list.append(f_tot)
    t_total = torch.stack(t_list, 1)
    f_total = torch.stack(f_list, 1)
    w_parent = torch.cat([t_total + torch.cross(r_p, f_total, dim=-1), f_total], -1) * hp
    w_child = torch.cat([t_total + torch.cross(r_c, f_total, dim=-1), f_total], -1)
    out = -w_child
    return out.index_add(1, m.parent_idx, w_parent)


# --------------------------------------------------------------------------- K5: integrate
def integrate_bodies(m: OracleModel, body_q, body_qd, body_f, inv_m, I, inv_I, dt):
    """inv_m [bs,nb], I / inv_I [bs,nb,3,3]. ``body_mass`` is read but unused by the reference
    kernel (integrator_euler.py:43) so it is not an argument."""
    x0, r0 = body_q[..., :3], body_q[..., 3:7]
    w0, v0 = body_qd[..., :3], body_qd[..., 3:]
    t0, f0 = body_f[..., :3], body_f[..., 3:]
    com = m.body_com[None]
    x_com = x0 + quat_rotate(r0, com)
    nz = (inv_m != 0).to(body_q.dtype)[..., None]
    v1 = v0 + (f0 * inv_m[..., None] + m.gravity * nz) * dt
    x1 = x_com + v1 * dt
    wb = quat_rotate_inv(r0, w0)
    Iwb = (I @ wb[..., None])[..., 0]
    tb = quat_rotate_inv(r0, t0) - torch.cross(wb, Iwb, dim=-1)
    w1 = quat_rotate(r0, wb + (inv_I @ tb[..., None])[..., 0] * dt)
    wq = torch.cat([w1, torch.zeros_like(w1[..., :1])], -1)
    r1 = safe_normalize(r0 + quat_mul(wq, r0) * 0.5 * dt)
    w1 = w1 * (1.0 - 0.1 * dt)
    w1 = wp_clamp(w1, -10.0, 10.0)
    v1 = wp_clamp(v1, -10.0, 10.0)
    q_new = torch.cat([x1 - quat_rotate(r1, com), r1], -1)
    qd_new = torch.cat([w1, v1], -1)
    return q_new, qd_new


# --------------------------------------------------------------------------- one substep / rollout
def substep(m, body_q, body_qd, res_f_t, refs_t, act_t, ke, kd, inv_m, I, inv_I, dt):
    """One ``simulate`` call (dp_model.py:1210-1228). Returns new state and the grf / jaf side channels
    (integrator_euler.py:510,544-548)."""
    f = res_f_t
    f = f - eval_body_contacts(m, body_q, body_qd)
    grf = f
    jf = eval_body_joints(m, body_q, body_qd, refs_t, act_t, ke, kd)
    f = f + jf
    q1, qd1 = integrate_bodies(m, body_q, body_qd, f, inv_m, I, inv_I, dt)
    return q1, qd1, grf.detach(), jf.detach()


def rollout(m: OracleModel, q_init, qd_init, torques, res_f, refs, target_ke, target_kd, body_inv_mass,
            body_inertia, body_inv_inertia, dt, frame_stride, num_frames, last_extra_step=True):
    """ForwardWarp.forward restated (dp_model.py:1147-1249) with [bs,...]-shaped tensors:
      q_init [bs,nq], qd_init [bs,nqd], torques / refs [T,bs,nqd], res_f [T,bs,nb,6],
      target_ke/kd [bs,nqd], body_inv_mass [bs,nb], body_inertia / body_inv_inertia [bs,nb,3,3].
    T = frame_stride*(num_frames-1)+1 substeps; outputs are states at t = k*frame_stride.
    Returns pos [F,bs,nb,7], vel [F,bs,nb,6], grfs [F,bs,nb,6], jafs [F,bs,nb,6]."""
    T = frame_stride * (num_frames - 1) + 1
    body_q, body_qd = eval_fk(m, q_init, qd_init)
    pos, vel, grfs, jafs = [], [], [], []
    for t in range(T):
        if t % frame_stride == 0:
            pos.append(body_q)
            vel.append(body_qd)
        if t == T - 1 and not last_extra_step:
            break
        q1, qd1, grf, jaf = substep(m, body_q, body_qd, res_f[t], refs[t], torques[t], target_ke, target_kd,
                                    body_inv_mass, body_inertia, body_inv_inertia, dt)
        if t % frame_stride == 0:
            grfs.append(grf)
            jafs.append(jaf)
        body_q, body_qd = q1, qd1
    out = (torch.stack(pos, 0), torch.stack(vel, 0))
    if grfs:
        out = out + (torch.stack(grfs, 0), torch.stack(jafs, 0))
    return out
