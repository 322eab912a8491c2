"""Motion-imitation optimisation model on top of the B200 rollout ops -- the caller of the hot path.

A compact host-side mirror of the reference's ``phys_model`` (/root/reference/diffphys/dp_model.py:56-1011): same
method names and data flow for the part that feeds and consumes the simulator boundary,

    preset_data (:407-427)  reinit_envs (:354-405)  compute_frame_start (:581-586)  get_batch_input (:611-662)
    forward (:664-838)      backward (:840)         update (:511-516)

and the same learnable quantities (control-reference networks, PD gains, body mass, global SE(3), initial velocity).
What is deliberately NOT mirrored (out of the hot-path scope): visualiser / ``query()`` mesh
articulation, lab4d coupling, checkpoint files on disk (the in-memory queue and its roll-back are mirrored:
save_checkpoint / update, and so is the per-parameter median gradient clipping: median_clip_).
Differences that matter for speed: mocap interpolation runs in torch on the device (the reference calls scipy on
the CPU every step, :605-609); shared parameters are passed UN-replicated to ``ForwardWarp`` (the reference
replicates + inverts bs*nb 3x3 matrices per step, :723-730).
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.nn as nn

from .model import ASSET_DIR, load_robot
from .ops import ForwardKinematics, ForwardWarp, ForwardWarpLoss, FrameCompose, Se3Loss, SimEnv, convert_ppr_warp


# ----------------------------------------------------------------------------------------- geometry (torch, xyzw)
def quat_to_matrix(q):
    x, y, z, w = q.unbind(-1)
    n = (q * q).sum(-1)
    s = 2.0 / n
    return torch.stack([1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w),
                        s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w),
                        s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)], -1).reshape(q.shape[:-1] + (3, 3))


def quat_mul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + bw * ax + ay * bz - az * by, aw * by + bw * ay + az * bx - ax * bz,
                        aw * bz + bw * az + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], -1)


def axis_angle_to_quat(v):
    ang = v.norm(dim=-1, keepdim=True)
    half = 0.5 * ang
    k = torch.where(ang > 1e-6, torch.sin(half) / ang.clamp_min(1e-6), 0.5 - ang * ang / 48.0)
    return torch.cat([v * k, torch.cos(half)], -1)


def rot_angle(mat, eps=1e-4):
    """geom_utils.py:37-46"""
    cos = (mat[..., 0, 0] + mat[..., 1, 1] + mat[..., 2, 2] - 1) / 2
    return torch.acos(cos.clamp(-1 + eps, 1 - eps))


def se3_loss_torch(pred, gt, rot_ratio=0.1):
    """dp_utils.py:113-138 for (...,7) poses (xyzw) or (...,6) twists, composed from torch ops (any device / dtype):
    the definition the fused kernel is tested against."""
    nan = torch.logical_or(pred.sum(-1).isnan(), gt.sum(-1).isnan())
    trn = (pred[..., :3] - gt[..., :3]).pow(2).sum(-1)
    if pred.shape[-1] == 7:
        rot = rot_angle(quat_to_matrix(pred[..., 3:]) @ quat_to_matrix(gt[..., 3:]).transpose(-1, -2))
    else:
        rot = rot_angle(quat_to_matrix(axis_angle_to_quat(pred[..., 3:])) @
                        quat_to_matrix(axis_angle_to_quat(gt[..., 3:])).transpose(-1, -2))
    loss = trn + rot * rot_ratio
    return torch.where(nan, torch.zeros_like(loss), loss)


def se3_loss(pred, gt, rot_ratio=0.1):
    """dp_utils.py:113-138 through the fused kernel pair (ops.Se3Loss: one launch forward, one backward).  CUDA
    tensors only -- there is no CPU path; `se3_loss_torch` is the composed definition it is tested against."""
    if not pred.is_cuda:
        raise RuntimeError("se3_loss needs CUDA tensors (the composed torch definition is se3_loss_torch)")
    return Se3Loss.apply(pred, gt.to(pred.device), rot_ratio)


def reduce_loss(loss_seq, clip=False):
    """dp_utils.py:93-110 on a [bs, T] loss: optional trajectory clipping, then the mean over the entries > 0 (all
    entries when none is).  Written with masks -- no boolean indexing, no Python loop over envs, no host read -- so the
    step can be captured in a CUDA graph.

    clip (used for the trajectory loss, dp_model.py:779): the threshold is 10 x the median of the positive entries of
    env 0 (the reference computes it in the first loop iteration and keeps it for all envs, :98-100); every env is cut
    from its first entry above the threshold onwards (:101-103).  The masked mean is identical to the reference's
    because the se3 losses are >= 0 (no entry > 0  <=>  every entry is 0  <=>  both means are 0)."""
    if clip:
        first = loss_seq[0]
        th = torch.nanmedian(torch.where(first > 0, first, torch.full_like(first, float("nan")))) * 10
        over = (loss_seq > th).to(loss_seq.dtype)           # (a NaN threshold -- no positive entry -- clips nothing)
        cut = torch.cumsum(over, dim=1) > 0                 # at or after the first entry above the threshold
        loss_seq = torch.where(cut, torch.zeros_like(loss_seq), loss_seq)
    pos = loss_seq > 0
    return torch.where(pos, loss_seq, torch.zeros_like(loss_seq)).sum() / pos.sum().clamp_min(1)


def median_clip_(named_grads, queue, queue_length=10, scale=5.0, norms=None):
    """Per-parameter outlier clipping of dp_model.py:965-998: every parameter keeps a queue of its recent gradient norms;
    once the queue holds more than ``queue_length`` entries, a gradient whose norm exceeds ``scale`` x the median of the
    queue (without its newest entry) is scaled down to that median and NOT enqueued, any other norm replaces the oldest
    entry.  ``named_grads``: iterable of (name, grad tensor); ``queue``: dict name -> list of floats, updated in place.
    All norms come to the host in ONE read (or are passed in as ``norms``, one float per non-None gradient).
    Returns {name: (norm, median or None, clipped)}."""
    named_grads = [(n, g) for n, g in named_grads if g is not None]
    if not named_grads:
        return {}
    if norms is None:
        norms = torch.stack(torch._foreach_norm([g for _, g in named_grads])).tolist()   # one multi-tensor launch, one read
    out = {}
    for (name, g), norm in zip(named_grads, norms):
        q = queue.setdefault(name, [])
        med, clipped = None, False
        if len(q) > queue_length:
            med = float(torch.tensor(q[:-1]).median())
            if norm > scale * med:
                g.mul_(med / (norm + 1e-6))                  # clip_grad_norm_(p, med_grad)
                clipped = True
            else:
                q.append(norm)
                q.pop(0)
        else:
            q.append(norm)
        out[name] = (norm, med, clipped)
    return out


def median_clip_device_(grads, queue, count, skip, queue_length=10, scale=5.0, norms=None):
    """``median_clip_`` as tensor operations only -- no host read, CUDA-graph capturable.  ``queue``: [n, queue_length+1]
    tensor of recent norms per gradient, oldest first; ``count``: 0-dim long tensor, entries held (all gradients enqueue in
    lock step until the queue is full); ``skip``: 0-dim bool tensor, True = the whole iteration is dropped (nothing is
    clipped or enqueued).  Scales the clipped gradients in place, updates ``queue`` / ``count`` in place and returns
    (norms [n], clipped [n] bool)."""
    n, L = len(grads), queue_length
    if norms is None:
        norms = torch.stack(torch._foreach_norm(grads))
    full = count > L
    med = queue[:, :L].median(dim=1).values              # lower median, like torch.tensor(q[:-1]).median()
    clipped = full & (norms > scale * med) & ~skip
    factor = torch.where(clipped, med / (norms + 1e-6), torch.ones_like(norms))
    torch._foreach_mul_(grads, list(factor.unbind()))
    enqueue = ~skip & ~clipped
    shifted = torch.cat([queue[:, 1:], norms[:, None]], 1)                       # full: drop the oldest
    filled = queue.scatter(1, count.clamp(max=L).reshape(1, 1).expand(n, 1), norms[:, None])
    queue.copy_(torch.where(enqueue[:, None], torch.where(full, shifted, filled), queue))
    count.add_((~skip & ~full).to(count.dtype))
    return norms, clipped


def rotate_frame(global_q, q):
    """T = T_global @ T  (dp_utils.py:60-72) on (...,7) poses."""
    R = quat_to_matrix(global_q[3:7])
    pos = q[..., :3] @ R.T + global_q[:3]
    return torch.cat([pos, quat_mul(global_q[3:7].expand_as(q[..., 3:7]), q[..., 3:7])], -1)


def rotate_frame_vel(global_q, qd):
    R = quat_to_matrix(global_q[3:7])
    return torch.cat([qd[..., :3] @ R.T, qd[..., 3:6] @ R.T], -1)


def compose_delta(q, delta):
    """dp_utils.py:21-30: T = T_delta @ T with delta = (translation, axis-angle)."""
    dq = axis_angle_to_quat(delta[..., 3:6])
    R = quat_to_matrix(dq)
    pos = (R @ q[..., :3, None])[..., 0] + delta[..., :3]
    return torch.cat([pos, quat_mul(dq, q[..., 3:7])], -1)


class TimeMLP(nn.Module):
    """Scalar-time -> vector network (Fourier time embedding + MLP + scaled head); stands in for the reference's
    TimeMLPWrapper (torch_utils.py:116-190). The last layer starts at zero so every predicted delta starts at 0."""

    def __init__(self, num_frames, out_channels, num_freq=6, width=128, depth=3, time_scale=1.0, output_scale=1.0):
        super().__init__()
        self.num_frames, self.time_scale, self.output_scale = num_frames, time_scale, output_scale
        self.register_buffer("freqs", 2.0 ** torch.arange(num_freq, dtype=torch.float32) * math.pi)
        layers, d = [], 2 * num_freq + 1
        for _ in range(depth):
            layers += [nn.Linear(d, width), nn.ReLU(True)]
            d = width
        self.body = nn.Sequential(*layers)
        self.head = nn.Linear(width, out_channels)
        nn.init.zeros_(self.head.weight)
        nn.init.zeros_(self.head.bias)

    def forward(self, frame_id):
        t = (frame_id.float() / self.num_frames * 2 - 1)[..., None] * self.time_scale
        emb = torch.cat([t, torch.sin(t * self.freqs), torch.cos(t * self.freqs)], -1)
        return self.head(self.body(emb)) * self.output_scale


def load_motion(seqname):
    z = np.load(os.path.join(ASSET_DIR, "motion_%s.npz" % seqname))
    return z["frames"].astype(np.float32), float(z["frame_duration"])


def parse_amp(amp):
    """dataloader.py:21-31 column map."""
    return dict(pos=amp[..., 0:3], orn=amp[..., 3:7], vel=amp[..., 31:34], avel=amp[..., 34:37],
                jang=amp[..., 7:19], jvel=amp[..., 37:49])


_BULLET2GL = torch.tensor([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]])


class ImitationModel(nn.Module):
    def __init__(self, robot="laikago", seqname="mi-trot", dt=5e-4, device="cuda", total_iters=101, lr=1e-4,
                 traj_wt=0.01, pos_state_wt=0.01, vel_state_wt=1e-4, noise_std=2e-3, seed=0, fused_traj_loss=True):
        super().__init__()
        # fused_traj_loss: the trajectory loss se3_loss(sim, target) (dp_model.py:777) is evaluated INSIDE the rollout
        # kernels (ops.ForwardWarpLoss) instead of by a separate kernel pair on the frame poses
        self.fused_traj_loss = bool(fused_traj_loss)
        self.device = torch.device(device)
        self.dt, self.noise_std = dt, noise_std
        self.wts = dict(traj=traj_wt, pos_state=pos_state_wt, vel_state=vel_state_wt)
        self.rng = np.random.RandomState(seed)
        self.robot_model = load_robot(robot)
        rm = self.robot_model
        self.n_dof, self.n_links = rm.nqd - 6, rm.nb
        frames, self.frame_interval = load_motion(seqname)
        self.preset_data(frames)
        self.env = SimEnv(rm, device=self.device)
        self.target_ke = nn.Parameter(torch.as_tensor(rm.joint_target_ke))
        self.target_kd = nn.Parameter(torch.as_tensor(rm.joint_target_kd))
        self.body_mass = nn.Parameter(torch.as_tensor(rm.body_mass))
        self.register_buffer("norm_body_inertia", torch.as_tensor(rm.norm_body_inertia))
        self.register_buffer("norm_body_inertia_inv", torch.linalg.inv(torch.as_tensor(rm.norm_body_inertia)))
        N = self.total_frames
        self.root_pose_mlp = TimeMLP(N, 6, time_scale=0.1, output_scale=0.5)
        self.joint_angle_mlp = TimeMLP(N, self.n_dof)
        self.vel_mlp = TimeMLP(N, 6 + self.n_dof, output_scale=5.0)
        self.global_q = nn.Parameter(torch.tensor([0.0, 0, 0, 0, 0, 0, 1.0]))
        self.to(self.device)
        self.init_global_q()
        self.progress = 0.0
        explicit = [self.global_q, self.target_ke, self.target_kd, self.body_mass]
        nets = [p for m in (self.root_pose_mlp, self.joint_angle_mlp, self.vel_mlp) for p in m.parameters()]
        # fused=True: one multi-tensor kernel per parameter group instead of ~10 foreach launches.  On CUDA the step is
        # capturable (learning rates live in device tensors the scheduler fills in place; a dropped iteration is the
        # optimizer's device-side ``found_inf`` flag), so that GraphedStep can replay it with the rest of the iteration
        cuda = self.device.type == "cuda"
        as_lr = (lambda v: torch.tensor(v, device=self.device)) if cuda else (lambda v: v)
        self.optimizer = torch.optim.AdamW([{"params": explicit, "lr": as_lr(lr * 10)}, {"params": nets, "lr": as_lr(lr)}],
                                           weight_decay=1e-4, fused=True, capturable=cuda)
        total = max(2, total_iters)
        self.scheduler = torch.optim.lr_scheduler.OneCycleLR(self.optimizer, [lr * 10, lr], total, pct_start=2.0 / total,
                                                             cycle_momentum=False, anneal_strategy="linear",
                                                             final_div_factor=1e2, div_factor=25)

    # ---- data ----------------------------------------------------------------------------------------
    def preset_data(self, frames):
        self.total_frames = frames.shape[0]
        self.steps_per_fr_interval = int(self.frame_interval / self.dt)
        self.register_buffer("amp_info", torch.as_tensor(frames))
        self.register_buffer("bullet2gl", _BULLET2GL.clone(), persistent=False)
        gl = torch.as_tensor(frames).clone()
        P = _BULLET2GL
        for a, b in ((0, 3), (3, 6), (31, 34), (34, 37)):     # pos, orn.xyz, vel, avel (parse_amp's column map)
            gl[:, a:b] = gl[:, a:b] @ P.T
        self.register_buffer("amp_gl", gl, persistent=False)

    def get_mocap_data(self, steps_fr):
        """linear interpolation / extrapolation of all columns at fractional frame ids (dp_model.py:421-427),
        then bullet2gl (dp_utils.py:141-156, in_bullet = False).  bullet2gl is linear and the same for every frame,
        so it is applied ONCE to the clip (`amp_gl`, preset_data) and the per-iteration work is one gather + lerp."""
        f0 = steps_fr.floor().clamp(0, self.total_frames - 2).long()
        w = (steps_fr - f0.float())[..., None]
        amp = torch.lerp(self.amp_gl[f0], self.amp_gl[f0 + 1], w)
        return parse_amp(amp)

    def lowest_point(self, body_q):
        """min world-y over the collision vertices (stands in for get_foot_height's mesh query, :574-579)."""
        rm = self.robot_model
        cb = torch.as_tensor(rm.contact_body, dtype=torch.long, device=body_q.device)
        cp = torch.as_tensor(rm.contact_point, device=body_q.device)
        R = quat_to_matrix(body_q[..., cb, 3:7])
        y = (R @ cp[:, :, None])[..., 1, 0] + body_q[..., cb, 1]
        return y.min(-1)[0]

    @torch.no_grad()
    def init_global_q(self):
        """ground-align the clip: lowest collision vertex of frame 0 touches y = 0 (dp_model.py:243-267)."""
        m = self.get_mocap_data(torch.zeros(1, 1, device=self.device))
        q = torch.cat([m["pos"], m["orn"], m["jang"]], -1)  # 1,1,nq
        qd = torch.zeros(1, 1, self.env.nqd, device=self.device)
        bq, _, _ = ForwardKinematics.apply(q, qd, self.env)
        self.global_q.data[1] = -self.lowest_point(bq)[0, 0]

    def reinit_envs(self, num_envs, frames_per_wdw, is_eval=False):
        self.num_envs, self.frames_per_wdw = num_envs, frames_per_wdw
        spf = self.steps_per_fr_interval
        self.steps_idx = range(spf * (frames_per_wdw - 1) + 1)
        self.steps_idx_fr = torch.arange(len(self.steps_idx), device=self.device).float() / spf
        self.frame2step = [i for i in self.steps_idx if i % spf == 0]
        self.is_eval = is_eval

    def compute_frame_start_host(self):
        fs = self.rng.rand(self.num_envs) * (self.total_frames - self.frames_per_wdw)
        return torch.as_tensor(np.round(fs), dtype=torch.float32)

    def compute_frame_start(self):
        return self.compute_frame_start_host().to(self.device)

    def fk_pos_vel(self, q, ja, qd=None, jad=None):
        """(bs,F,..) targets -> body poses / twists through ForwardKinematics (dp_model.py:588-603).  ``qd`` None: poses
        only (the twists come back zero)."""
        tq = torch.cat([q, ja], -1).permute(1, 0, 2).contiguous()
        tqd = None if qd is None else convert_ppr_warp(torch.cat([qd, jad], -1).permute(1, 0, 2).contiguous())
        bq, bqd, frames = ForwardKinematics.apply(tq, tqd, self.env)
        return bq, convert_ppr_warp(bqd), frames

    def get_batch_input(self, steps_fr):
        msm = self.get_mocap_data(steps_fr)
        fid = steps_fr.reshape(-1)
        bs, T = steps_fr.shape
        delta_root = self.root_pose_mlp(fid).view(bs, T, 6)
        # rotate_frame(global_q, .) then compose_delta(., delta_root): one fused kernel each way (ops.FrameCompose)
        target_q, queried_q = FrameCompose.apply(self.global_q, torch.cat([msm["pos"], msm["orn"]], -1), delta_root)
        # the reference also rotates the mocap twists into the target FK (rotate_frame_vel, dp_model.py:626-636), whose
        # twist OUTPUT (target_velocity) nothing reads: not computed here
        f2s = slice(0, None, self.steps_per_fr_interval)   # == self.frame2step (evenly strided), as a view: no index tensor
        target_position, _, self.target_trajs = self.fk_pos_vel(target_q[:, f2s], msm["jang"][:, f2s])
        delta_ja = self.joint_angle_mlp(fid).view(bs, T, -1)
        queried_qd = self.vel_mlp(fid).view(bs, T, -1)
        queried_ja = msm["jang"] + delta_ja
        # time-major flattening (rearrange_pred, :555-572)
        q_all = torch.cat([queried_q, queried_ja], -1).permute(1, 0, 2).reshape(T, -1)
        qd_all = queried_qd.permute(1, 0, 2).reshape(T, -1)
        ref_ja = torch.cat([torch.zeros_like(queried_ja[..., :6]), queried_ja], -1).permute(1, 0, 2).reshape(T, -1)
        return target_position, ref_ja, q_all, qd_all

    # ---- one optimisation step -------------------------------------------------------------------------
    def draw_noise(self):
        """initial-state noise of a training iteration (dp_model.py:700-712), drawn on the host like the reference;
        None when the iteration adds none (eval)."""
        if not (self.training and self.noise_std > 0 and not self.is_eval):
            return None
        ratio = float(np.clip(1 - 1.5 * self.progress, 0, 1))
        noise = torch.as_tensor(self.rng.normal(size=(self.num_envs, self.env.nq), scale=self.noise_std * ratio),
                                dtype=torch.float32)
        noise[:, :3] = 0
        noise[:, 3:7] *= 5
        return noise

    def forward(self, frame_start=None, noise=None):
        """``noise``: [bs, nq] tensor added to q_init (GraphedStep passes a static buffer); None = draw it here."""
        if frame_start is None:
            frame_start = self.compute_frame_start()
        if noise is None:
            noise = self.draw_noise()
            if noise is not None:
                noise = noise.to(self.device)
        steps_fr = frame_start[:, None] + self.steps_idx_fr[None]                     # bs,T
        target_position, ref_ja, queried_q, queried_qd = self.get_batch_input(steps_fr)
        bs, F = self.num_envs, self.frames_per_wdw
        q_init = queried_q[0].reshape(-1)
        if noise is not None:
            q_init = q_init + noise.reshape(-1)
        qd_init = convert_ppr_warp(queried_qd[0].view(bs, -1)).reshape(-1)
        inv_m = 1.0 / self.body_mass
        I = self.norm_body_inertia * self.body_mass[:, None, None]
        inv_I = self.norm_body_inertia_inv * inv_m[:, None, None]   # inverse(nI * m) = inverse(nI) / m
        traj = None
        if self.fused_traj_loss:
            tgt = target_position.permute(1, 0, 2, 3).reshape(F, -1, 7)        # frame-major rows, like wp_pos
            loss_pos, sim_position, sim_velocity = ForwardWarpLoss.apply(
                q_init, qd_init, None, None, ref_ja, self.target_ke, self.target_kd, self.body_mass, inv_m, I, inv_I, tgt,
                0.1, self)
            traj = loss_pos.view(F, bs, -1).permute(1, 0, 2).mean(-1)          # [bs, F], mean over bodies
        else:
            sim_position, sim_velocity = ForwardWarp.apply(q_init, qd_init, None, None, ref_ja, self.target_ke,
                                                           self.target_kd, self.body_mass, inv_m, I, inv_I, self)
        sim_velocity = convert_ppr_warp(sim_velocity)
        f2s = slice(0, None, self.steps_per_fr_interval)
        qq = queried_q[f2s].reshape(F, bs, -1)
        qqd = convert_ppr_warp(queried_qd[f2s].reshape(F, bs, -1))
        queried_position, queried_velocity, self.pid_ref = ForwardKinematics.apply(qq, qqd, self.env)
        queried_velocity = convert_ppr_warp(queried_velocity)
        sim_position = sim_position.reshape(F, bs, -1, 7).permute(1, 0, 2, 3)
        sim_velocity = sim_velocity.reshape(F, bs, -1, 6).permute(1, 0, 2, 3)
        loss_dict = {
            "traj": reduce_loss(traj if traj is not None else se3_loss(sim_position, target_position).mean(-1), clip=True),
            "pos_state": reduce_loss(se3_loss(queried_position, sim_position.detach()).mean(-1)),
            "vel_state": reduce_loss(se3_loss(queried_velocity, sim_velocity.detach()).mean(-1)),
        }
        total = sum(v * self.wts[k] for k, v in loss_dict.items())
        out = {"loss_" + k: v for k, v in loss_dict.items()}
        out["total_loss"] = total
        return out

    def backward(self, loss):
        loss.backward()

    def save_checkpoint(self):
        """In-memory queue of length two of (model, optimizer, scheduler) states (dp_model.py:912-921; the reference
        also writes ckpt_phys_%04d.pth, which is out of scope here).  main.py calls it every `iters_per_round` iterations."""
        from copy import deepcopy
        if not hasattr(self, "_cache"):
            self._cache = [None, None]
        self._cache[0] = self._cache[1]
        self._cache[1] = (deepcopy(self.state_dict()), deepcopy(self.optimizer.state_dict()),
                          deepcopy(self.scheduler.state_dict()))

    def update_device(self, thresh=10.0):
        """The device part of ``update``: gradient norms, the drop decision, per-parameter median clipping and the AdamW
        step as tensor operations with NO host read (CUDA-graph capturable).  ``thresh``: float or 0-dim tensor.  A
        dropped iteration (norm non-finite or above ``thresh``) is the optimizer's ``found_inf`` flag: the fused AdamW
        kernel leaves parameters, moments and step counts untouched.  Leaves [grad_norm, dropped] in ``self._upd_info``."""
        named = [(n, p) for n, p in self.named_parameters() if p.grad is not None]
        grads = [p.grad for _, p in named]
        if not hasattr(self, "_clip_queue"):
            dev = grads[0].device
            self._clip_names = [n for n, _ in named]
            self._clip_queue = torch.zeros(len(named), 11, device=dev)
            self._clip_count = torch.zeros((), dtype=torch.long, device=dev)
            self._upd_info = torch.zeros(2, device=dev)
            self.optimizer.found_inf = torch.zeros((), device=dev)
        assert [n for n, _ in named] == self._clip_names, "the set of parameters with gradients changed"
        norms = torch.stack(torch._foreach_norm(grads))       # every per-parameter norm in one multi-tensor launch
        total = norms.norm(2)
        skip = ~(total <= thresh)                             # NaN / inf / above the threshold (dp_model.py:941-946)
        median_clip_device_(grads, self._clip_queue, self._clip_count, skip, norms=norms)   # dp_model.py:965-998
        self.optimizer.found_inf.copy_(skip)
        self.optimizer.step()
        self._upd_info.copy_(torch.stack([total, skip.to(total.dtype)]))

    def finish_update(self, keep_grads=False):
        """The host part of ``update``: ONE read of [grad_norm, dropped]; a dropped iteration rolls the model and the
        optimizer back IN PLACE (parameter / moment storage, and with it a captured graph, stays valid) to the snapshot
        of two rounds ago when there is one (dp_model.py:947-952); the scheduler steps either way."""
        from copy import deepcopy
        grad_norm, skipped = self._upd_info.tolist()
        skipped = bool(skipped)
        if skipped and getattr(self, "_cache", [None])[0] is not None:
            sd, od, sch = self._cache[0]
            with torch.no_grad():
                for k, v in self.state_dict().items():
                    v.copy_(sd[k])
                params = [p for g in self.optimizer.param_groups for p in g["params"]]
                for i, p in enumerate(params):
                    for k, v in self.optimizer.state.get(p, {}).items():
                        saved = od["state"].get(i, {}).get(k)
                        v.zero_() if saved is None else v.copy_(saved)   # (snapshot older than the first step: zeros)
            self.scheduler.load_state_dict(deepcopy(sch))    # (the learning-rate tensors are re-filled by scheduler.step() below)
        self.scheduler.step()
        if not keep_grads:
            self.optimizer.zero_grad()
        return {"grad_norm": float(grad_norm), "skipped": skipped}

    def update(self, thresh=10.0, keep_grads=False):
        """clip, sanity-check and apply the gradients (dp_model.py:511-548,936-963): an iteration whose gradient norm
        is non-finite or above ``thresh`` is skipped (the reference drops the gradients, which makes its optimizer step
        a no-op) and the model / optimizer / scheduler roll back to the snapshot of two rounds ago when there is one
        (:947-952).  ``keep_grads``: leave the .grad tensors in place."""
        self.update_device(thresh)
        return self.finish_update(keep_grads)


class GraphedStep:
    """One whole optimisation iteration -- forward, losses, backward, gradient norms, drop decision, median clipping and
    the AdamW step -- captured ONCE in a CUDA graph and replayed.

    The reference-shaped problem (10..64 windows x 760 substeps, dp_model.py:354-367) cannot fill a B200: an iteration
    is a few hundred small launches (three MLPs, mocap interpolation, two FK calls, the rollout pair, se3 losses
    and their backward, the optimizer) whose launch overhead, not their run time, sets the iteration time.  Everything
    that changes between iterations enters through static device buffers (window start frames, initial-state noise, the
    drop threshold, the learning rates), the first two drawn on the host exactly like the eager path, so the two paths
    consume the same random stream.  The host reads two floats per iteration (gradient norm, dropped flag) and keeps the
    rare roll-back and the scheduler (``ImitationModel.finish_update``).  ``capture_update=False`` captures forward +
    backward only and runs ``update()`` eagerly.

    Limitation: the kernel arguments are baked into the captured graph BY VALUE, including the model scalars (attach
    gains, gravity, ground flag, checkpoint policy) and the joint_X_p pointer -- ``env.set_attach / set_gravity /
    ground / joint_X_p = other_tensor`` after the capture are NOT seen by replays (in-place updates of a per-env
    joint_X_p tensor are).  ``env._snapshot()`` is recorded at capture time and checked on every replay.

        step = GraphedStep(model)           # after model.train(); model.reinit_envs(...)
        for it in range(iters):
            model.progress = it / (iters - 1)
            out, info = step()              # out: dict of static loss tensors, info: update() result
    """

    def __init__(self, model, warmup=3, capture_update=True):
        self.model = m = model
        dev = m.device
        self.capture_update = bool(capture_update)
        self._snap = m.env._snapshot()[:-1]      # (the tensor version legitimately changes under in-place updates)
        self.frame_start = torch.zeros(m.num_envs, device=dev)
        self.noise = torch.zeros(m.num_envs, m.env.nq, device=dev)
        self.thresh = torch.full((), -1.0, device=dev)   # -1: every warm-up iteration is "dropped" -> nothing changes
        self._thresh_host = -1.0
        self._h_frame_start = torch.zeros(m.num_envs).pin_memory()
        self._h_noise = torch.zeros(m.num_envs, m.env.nq).pin_memory()
        m.optimizer.zero_grad(set_to_none=True)

        def iteration():
            out = m(frame_start=self.frame_start, noise=self.noise)
            out["total_loss"].backward()
            if self.capture_update:
                m.update_device(self.thresh)
            return out
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):       # warm-up off the default stream (allocator pools, lazy kernel attributes,
            for _ in range(warmup):         # the optimizer's lazily created moments)
                iteration()
                m.optimizer.zero_grad(set_to_none=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = iteration()
        self.launches_per_replay = None

    def __call__(self, frame_start=None, thresh=10.0):
        m = self.model
        if m.env._snapshot()[:-1] != self._snap:
            raise RuntimeError("SimEnv was modified after the CUDA graph was captured; re-create the GraphedStep")
        if frame_start is None:
            self._h_frame_start.copy_(m.compute_frame_start_host())
            self.frame_start.copy_(self._h_frame_start, non_blocking=True)
        else:
            self.frame_start.copy_(torch.as_tensor(frame_start, dtype=torch.float32), non_blocking=True)
        noise = m.draw_noise()
        if noise is None:
            self.noise.zero_()
        else:
            self._h_noise.copy_(noise)
            self.noise.copy_(self._h_noise, non_blocking=True)
        if self.capture_update and float(thresh) != self._thresh_host:
            self._thresh_host = float(thresh)
            self.thresh.fill_(self._thresh_host)
        self.graph.replay()
        if self.capture_update:
            return self.out, m.finish_update(keep_grads=True)
        return self.out, m.update(thresh, keep_grads=True)
