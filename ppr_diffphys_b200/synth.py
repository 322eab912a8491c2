"""Synthetic batched rollout inputs on the device: root at standing height,
identity orientation + N(0,0.01) quaternion noise (renormalised), joint angles U(-a,a), qd_init N(0,0.1),
refs = joint angles + 0.1 sin(2 pi t / 64 + phi_j) -- sampled per substep, or (``frame_stride``) sampled at the frame
steps and linearly interpolated in between like the reference's mocap targets (dp_model.py:421-427): then
``ref_frames`` [F, bs*nqd] is what a caller ships and ``refs`` its expansion; masses / inertias / PD gains from the
compiled robot.
Returned tensors use the reference's flattened call layout of ForwardWarp.apply (dp_model.py:563-572,697-699)."""
from __future__ import annotations

import math

import torch


def _quat_rotate(q, v):
    u, w = q[..., :3], q[..., 3:]
    return v * (2.0 * w * w - 1.0) + 2.0 * w * torch.cross(u, v, dim=-1) + 2.0 * u * (u * v).sum(-1, keepdim=True)


def lerp_frames(frames, stride, nsteps):
    """[F, n] per-frame values -> [nsteps, n], linear in between (same arithmetic as ppr_refs_from_frames)."""
    t = torch.arange(nsteps)
    k0 = torch.clamp(t // stride, max=max(frames.shape[0] - 2, 0))
    a = ((t - k0 * stride).float() / float(stride))[:, None]
    k1 = torch.clamp(k0 + 1, max=frames.shape[0] - 1)
    v0, v1 = frames[k0], frames[k1]
    return (1.0 - a) * v0 + a * v1


def make_batch(env, bs, nsteps, seed=0, clearance=1e-3, ang=0.2, qd_std=0.1, ref_amp=0.1, lin_vel=0.0,
               pinned_host=False, frame_stride=None):
    """env: SimEnv. clearance > 0: lowest contact point that far ABOVE ground; < 0: penetrating."""
    rm, dev = env.model, env.device
    g = torch.Generator(device="cpu").manual_seed(seed)
    nb, nq, nqd = rm.nb, rm.nq, rm.nqd
    B = nqd - 6
    ja = (torch.rand(bs, B, generator=g) * 2 - 1) * ang
    quat = torch.tensor([0.0, 0, 0, 1.0]).expand(bs, 4) + torch.randn(bs, 4, generator=g) * 0.01
    quat = quat / quat.norm(dim=-1, keepdim=True)
    pos = torch.zeros(bs, 3)
    q_init = torch.cat([pos, quat, ja], -1)
    qd_init = torch.randn(bs, nqd, generator=g) * qd_std
    if lin_vel > 0:
        qd_init[:, 3] = (torch.rand(bs, generator=g) * 2 - 1) * lin_vel
        qd_init[:, 5] = (torch.rand(bs, generator=g) * 2 - 1) * lin_vel
    # settle the root height with the library's own FK
    bq, _ = env.fk(q_init.to(dev), torch.zeros(bs, nqd, device=dev))
    cb = torch.as_tensor(rm.contact_body, dtype=torch.long, device=dev)
    cp = torch.as_tensor(rm.contact_point, device=dev)
    cd = torch.as_tensor(rm.contact_dist, device=dev)
    X = bq[:, cb]
    y = (X[..., :3] + _quat_rotate(X[..., 3:7], cp[None]))[..., 1] - cd[None]
    q_init[:, 1] = (-y.min(dim=1)[0] + clearance).cpu()
    t = torch.arange(nsteps, dtype=torch.float32)[:, None, None]
    phase = torch.rand(1, bs, B, generator=g) * 2 * math.pi
    refs = torch.zeros(nsteps, bs, nqd)
    refs[:, :, 6:] = ja[None] + ref_amp * torch.sin(2 * math.pi * t / 64.0 + phase)
    ref_frames = None
    if frame_stride:
        nfr = (nsteps - 1 + frame_stride - 1) // frame_stride + 1
        tf = (torch.arange(nfr, dtype=torch.float32) * frame_stride)[:, None, None]
        ref_frames = torch.zeros(nfr, bs, nqd)
        ref_frames[:, :, 6:] = ja[None] + ref_amp * torch.sin(2 * math.pi * tf / 64.0 + phase)
        ref_frames = ref_frames.reshape(nfr, -1).contiguous()
        refs = lerp_frames(ref_frames, frame_stride, nsteps).reshape(nsteps, bs, nqd)
    ke = torch.as_tensor(rm.joint_target_ke)[None].repeat(bs, 1)
    kd = torch.as_tensor(rm.joint_target_kd)[None].repeat(bs, 1)
    mass = torch.as_tensor(rm.body_mass)[None].repeat(bs, 1)
    host = dict(q_init=q_init.reshape(-1), qd_init=qd_init.reshape(-1), refs=refs.reshape(nsteps, -1),
                target_ke=ke.reshape(-1), target_kd=kd.reshape(-1), body_mass=mass.reshape(-1))
    if ref_frames is not None:
        host["ref_frames"] = ref_frames
    host = {k: v.float().contiguous() for k, v in host.items()}
    if pinned_host:
        host = {k: v.pin_memory() for k, v in host.items()}
    return host


def shared_param_chain(target_ke, target_kd, body_mass, norm_body_inertia, bs):
    """Shared parameters [nqd], [nqd], [nb] -> the per-env replicated tensors ForwardWarp.apply expects
    (dp_model.py:723-730).  Same values as the reference's chain, but the 3x3 inverse is taken on the nb shared
    matrices BEFORE replication instead of on bs*nb copies of them; autograd sums the per-env gradients back."""
    nb = body_mass.numel()
    inv_m = 1.0 / body_mass
    I = norm_body_inertia * body_mass[:, None, None]
    inv_I = torch.linalg.inv(I)
    rep = lambda t: t[None].expand(bs, *t.shape).reshape(bs * t.shape[0], *t.shape[1:])
    return rep(target_ke), rep(target_kd), rep(body_mass), rep(inv_m), rep(I), rep(inv_I)
