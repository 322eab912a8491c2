"""Drop-in replacements of the reference's simulator boundary (diffphys/dp_model.py:1014-1400).

  convert_ppr_warp    dp_model.py:1014-1019
  ForwardKinematics   dp_model.py:1022-1130   (Warp eval_fk under wp.Tape  ->  ppr_fk_forward / ppr_fk_backward)
  ForwardWarp         dp_model.py:1145-1400   (per-substep Warp launches under wp.Tape  ->  ONE persistent rollout
                                               kernel + ONE hand-written adjoint kernel)
  SimEnv              the ``env`` object (reference: Warp ``Model`` built in reinit_envs, dp_model.py:384-401)

Same call signatures, argument meaning, output shapes and side effects (``self.grfs``, ``self.jafs``,
``self.sim_trajs``).  All device work goes through the C ABI in include/ppr_b200.h with raw pointers of torch
tensors and torch's current stream; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._capi import RolloutIO, make_desc
from .model import RobotModel, load_robot


def convert_ppr_warp(tensor):
    """[linear, angular, ...] <-> [angular, linear, ...] (dp_model.py:1014-1019)."""
    return torch.cat([tensor[..., 3:6], tensor[..., 0:3], tensor[..., 6:]], -1)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t, device):
    if t is None:
        return None
    if t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class LazyFrames:
    """List-like view of per-frame numpy arrays (env 0) that copies to the host only when read; the reference
    eagerly does ``.numpy()`` per frame (dp_model.py:1072,1244), forcing a sync on every forward."""

    def __init__(self, tensor):  # [F, nb, 7] on device
        self._t = tensor
        self._np = None

    def _get(self):
        if self._np is None:
            self._np = self._t.detach().cpu().numpy()
        return self._np

    def __len__(self):
        return int(self._t.shape[0])

    def __getitem__(self, i):
        return self._get()[i]

    def __iter__(self):
        return iter(self._get())


class SimEnv:
    """Static model of one articulation shared by ``num_envs`` environments, resident on one GPU.

    Mirrors what the reference reads from the Warp ``Model``: ``joint_X_p`` (mutable -- lab4d overwrites it,
    dp_interface.py:465), ``joint_attach_ke/kd``, ``gravity``, ``ground`` and the body/joint/contact tables.
    The reference replicates the tables num_envs times (dp_model.py:384-386); here ONE copy is uploaded."""

    def __init__(self, robot, device=None):
        if not torch.cuda.is_available():
            raise _lib.PprError("ppr_diffphys_b200 needs a CUDA device; there is no CPU fallback")
        self.model: RobotModel = load_robot(robot) if isinstance(robot, str) else robot
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = _lib.lib()
        desc, keep = make_desc(self.model)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.ppr_model_create(C.byref(desc), C.byref(h)), "ppr_model_create")
        self._h = h
        self._ground = True
        self.checkpoint_every = 1
        self._gravity = tuple(float(x) for x in self.model.gravity)
        self.nb, self.nq, self.nqd = self.model.nb, self.model.nq, self.model.nqd
        self.body_count_per_env = self.nb
        self._attach = [float(self.model.joint_attach_ke), float(self.model.joint_attach_kd)]
        self.body_com = torch.as_tensor(self.model.body_com)
        self._joint_X_p = torch.as_tensor(self.model.joint_X_p).clone()

    # -- mutable model attributes -----------------------------------------------------------------
    @property
    def ground(self):
        return self._ground

    @ground.setter
    def ground(self, value):
        """``env.ground = False`` skips the ground contacts like the reference does (integrator_euler.py:492)."""
        self._ground = bool(value)
        _lib.check(self._lib.ppr_model_set_ground(self._h, int(self._ground)), "ppr_model_set_ground")

    def _snapshot(self):
        """Everything of the mutable model state that a rollout's forward and backward must agree on."""
        xp = self._joint_X_p
        return (self.checkpoint_every, self.latency_envs, self.team_envs, self.joint_attach_ke, self.joint_attach_kd, self._gravity,
                self._ground, xp.data_ptr() if xp.is_cuda else None, tuple(xp.shape), xp._version)

    @property
    def joint_X_p(self):
        return self._joint_X_p

    @joint_X_p.setter
    def joint_X_p(self, value):
        """``env.joint_X_p = tensor`` as lab4d does every step (dp_interface.py:465).  [nb,7]: one table shared by
        all envs (uploaded).  [n_env*nb,7]: one block per env, the reference's own shape (every video instance has
        its own bone lengths) -- a CUDA tensor is used in place, zero-copy like wp.from_torch, and kept alive here."""
        v = torch.as_tensor(value).detach().reshape(-1, 7)
        if v.shape[0] > self.nb:
            assert v.shape[0] % self.nb == 0, "joint_X_p must have a multiple of nb rows"
            v = v.to(self.device, torch.float32).contiguous()
            self._joint_X_p = v
            _lib.check(self._lib.ppr_model_set_joint_X_p_env(self._h, C.c_void_p(v.data_ptr()), v.shape[0] // self.nb),
                       "ppr_model_set_joint_X_p_env")
            return
        v = v.float().cpu()[: self.nb].contiguous()
        self._joint_X_p = v
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()  # the host staging buffer is reused by the library
            _lib.check(self._lib.ppr_model_set_joint_X_p_env(self._h, None, 0), "ppr_model_set_joint_X_p_env")
            _lib.check(self._lib.ppr_model_set_joint_X_p(self._h, C.c_void_p(v.data_ptr()), _stream()),
                       "ppr_model_set_joint_X_p")

    def set_attach(self, ke, kd):
        self._attach = [float(ke), float(kd)]
        _lib.check(self._lib.ppr_model_set_attach(self._h, C.c_float(ke), C.c_float(kd)), "ppr_model_set_attach")

    # ``env.joint_attach_ke = ...`` / ``env.joint_attach_kd = ...`` as the reference assigns them (dp_model.py:392-393)
    @property
    def joint_attach_ke(self):
        return self._attach[0]

    @joint_attach_ke.setter
    def joint_attach_ke(self, v):
        self.set_attach(float(v), self._attach[1])

    @property
    def joint_attach_kd(self):
        return self._attach[1]

    @joint_attach_kd.setter
    def joint_attach_kd(self, v):
        self.set_attach(self._attach[0], float(v))

    def set_gravity(self, g):
        arr = (C.c_float * 3)(*[float(x) for x in g])
        _lib.check(self._lib.ppr_model_set_gravity(self._h, arr), "ppr_model_set_gravity")
        self._gravity = tuple(float(x) for x in g)

    def set_checkpoint_every(self, every):
        """Checkpoint policy K of the rollout: keep the state every K substeps and let the adjoint recompute the rest
        (K = 1, the default, is fastest; K > 1 shrinks the workspace K-fold for ~ +40 % time)."""
        _lib.check(self._lib.ppr_model_set_checkpoint_every(self._h, int(every)), "ppr_model_set_checkpoint_every")
        self.checkpoint_every = int(every)

    def set_latency_envs(self, max_envs):
        """Batches of at most ``max_envs`` environments run one environment per warp (lowest substep latency)."""
        _lib.check(self._lib.ppr_model_set_latency_envs(self._h, int(max_envs)), "ppr_model_set_latency_envs")

    @property
    def latency_envs(self):
        return int(self._lib.ppr_model_latency_envs(self._h))

    def set_team_envs(self, max_envs):
        """Batches of at most ``max_envs`` environments run one environment per block of three warps, the ground contacts
        on two helper warps (team layout: lowest substep latency for a few hundred environments at most)."""
        _lib.check(self._lib.ppr_model_set_team_envs(self._h, int(max_envs)), "ppr_model_set_team_envs")

    @property
    def team_envs(self):
        return int(self._lib.ppr_model_team_envs(self._h))

    @property
    def packing(self):
        """(threads per group, environments per group): a group is a warp or a thread block."""
        return int(self._lib.ppr_model_group_threads(self._h)), int(self._lib.ppr_model_envs_per_group(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.ppr_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- raw (non-autograd) entry points ------------------------------------------------------------
    def fk(self, joint_q, joint_qd):
        """joint_q [n,nq], joint_qd [n,nqd] -> body_q [n,nb,7], body_qd [n,nb,6]."""
        n = joint_q.shape[0]
        bq = torch.empty(n, self.nb, 7, device=self.device, dtype=torch.float32)
        bqd = torch.empty(n, self.nb, 6, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.ppr_fk_forward(self._h, n, _ptr(joint_q), _ptr(joint_qd), _ptr(bq), _ptr(bqd),
                                                _stream()), "ppr_fk_forward")
        return bq, bqd

    def fk_backward(self, joint_q, joint_qd, adj_bq, adj_bqd):
        n = joint_q.shape[0]
        aq = torch.empty(n, self.nq, device=self.device, dtype=torch.float32)
        aqd = torch.empty(n, self.nqd, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.ppr_fk_backward(self._h, n, _ptr(joint_q), _ptr(joint_qd), _ptr(adj_bq),
                                                 _ptr(adj_bqd), _ptr(aq), _ptr(aqd), _stream()), "ppr_fk_backward")
        return aq, aqd

    def workspace_bytes(self, bs, nsteps):
        return int(self._lib.ppr_rollout_workspace_bytes(self._h, bs, nsteps))

    def rollout_forward(self, bs, nsteps, stride, dt, q_init, qd_init, torques, res_f, refs, ke, kd, inv_m, I, inv_I,
                        want_forces=True, workspace=None, shared_params=False):
        F = (nsteps - 1) // stride + 1
        dev = self.device
        pos = torch.empty(F, bs * self.nb, 7, device=dev, dtype=torch.float32)
        vel = torch.empty(F, bs * self.nb, 6, device=dev, dtype=torch.float32)
        grf = torch.empty(F, bs * self.nb, 6, device=dev, dtype=torch.float32) if want_forces else None
        jaf = torch.empty(F, bs * self.nb, 6, device=dev, dtype=torch.float32) if want_forces else None
        nbytes = self.workspace_bytes(bs, nsteps)
        if workspace is None or workspace.numel() * workspace.element_size() < nbytes:
            workspace = torch.empty((nbytes + 3) // 4, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(self._lib.ppr_rollout_forward(
                self._h, bs, nsteps, stride, C.c_float(dt), int(bool(shared_params)), _ptr(q_init), _ptr(qd_init),
                _ptr(torques), _ptr(res_f),
                _ptr(refs), _ptr(ke), _ptr(kd), _ptr(inv_m), _ptr(I), _ptr(inv_I), _ptr(pos), _ptr(vel), _ptr(grf),
                _ptr(jaf), _ptr(workspace), C.c_size_t(workspace.numel() * 4), _stream()), "ppr_rollout_forward")
        return pos, vel, grf, jaf, workspace

    def rollout_backward(self, bs, nsteps, stride, dt, q_init, qd_init, torques, res_f, refs, ke, kd, inv_m, I, inv_I,
                         adj_pos, adj_vel, workspace, shared_params=False):
        dev = self.device
        e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        g = dict(q_init=e(bs * self.nq), qd_init=e(bs * self.nqd),
                 torques=e(nsteps, bs * self.nqd) if torques is not None else None,
                 res_f=e(nsteps, bs * self.nb, 6) if res_f is not None else None,
                 refs=e(nsteps, bs * self.nqd), target_ke=e(bs * self.nqd), target_kd=e(bs * self.nqd),
                 body_inv_mass=e(bs * self.nb), body_inertia=e(bs * self.nb, 3, 3),
                 body_inv_inertia=e(bs * self.nb, 3, 3))
        with torch.cuda.device(dev):
            _lib.check(self._lib.ppr_rollout_backward(
                self._h, bs, nsteps, stride, C.c_float(dt), int(bool(shared_params)), _ptr(q_init), _ptr(qd_init),
                _ptr(torques), _ptr(res_f),
                _ptr(refs), _ptr(ke), _ptr(kd), _ptr(inv_m), _ptr(I), _ptr(inv_I), _ptr(adj_pos), _ptr(adj_vel),
                _ptr(g["q_init"]), _ptr(g["qd_init"]), _ptr(g["torques"]), _ptr(g["res_f"]), _ptr(g["refs"]),
                _ptr(g["target_ke"]), _ptr(g["target_kd"]), _ptr(g["body_inv_mass"]), _ptr(g["body_inertia"]),
                _ptr(g["body_inv_inertia"]), _ptr(workspace), C.c_size_t(workspace.numel() * 4), _stream()),
                "ppr_rollout_backward")
        return g


    def rollout_backward_shared(self, bs, nsteps, stride, dt, q_init, qd_init, torques, res_f, refs, ke, kd, inv_m, I, inv_I,
                                adj_pos, adj_vel, workspace):
        """Adjoint with UN-replicated parameters; their gradients come back summed over the environments (reduced in
        the adjoint kernel's epilogue + one small reduce kernel) as views of ONE packed buffer ``g["packed"]`` =
        [target_ke | target_kd | body_inv_mass | body_inertia | body_inv_inertia]."""
        dev, nb, nqd = self.device, self.nb, self.nqd
        e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        P = 2 * nqd + 19 * nb
        packed = e(P)
        scratch = e((int(self._lib.ppr_rollout_reduce_scratch_bytes(self._h, bs)) + 3) // 4)
        g = dict(q_init=e(bs * self.nq), qd_init=e(bs * nqd),
                 torques=e(nsteps, bs * nqd) if torques is not None else None,
                 res_f=e(nsteps, bs * nb, 6) if res_f is not None else None, refs=e(nsteps, bs * nqd), packed=packed,
                 target_ke=packed[:nqd], target_kd=packed[nqd:2 * nqd], body_inv_mass=packed[2 * nqd:2 * nqd + nb],
                 body_inertia=packed[2 * nqd + nb:2 * nqd + 10 * nb].view(nb, 3, 3),
                 body_inv_inertia=packed[2 * nqd + 10 * nb:].view(nb, 3, 3))
        with torch.cuda.device(dev):
            _lib.check(self._lib.ppr_rollout_backward_shared(
                self._h, bs, nsteps, stride, C.c_float(dt), _ptr(q_init), _ptr(qd_init), _ptr(torques), _ptr(res_f),
                _ptr(refs), _ptr(ke), _ptr(kd), _ptr(inv_m), _ptr(I), _ptr(inv_I), _ptr(adj_pos), _ptr(adj_vel),
                _ptr(g["q_init"]), _ptr(g["qd_init"]), _ptr(g["torques"]), _ptr(g["res_f"]), _ptr(g["refs"]),
                _ptr(packed), _ptr(scratch), C.c_size_t(scratch.numel() * 4), _ptr(workspace),
                C.c_size_t(workspace.numel() * 4), _stream()), "ppr_rollout_backward_shared")
        return g


class ForwardKinematics(torch.autograd.Function):
    """``ForwardKinematics.apply(rj_q[T,bs,7+B], rj_qd[T,bs,6+B], env) -> (body_q[bs,T,nb,7], body_qd[bs,T,nb,6],
    body_q_numpy)`` -- dp_model.py:1022-1130. One launch for all T frames (reference: T launches + T State allocs)."""

    @staticmethod
    def forward(ctx, rj_q, rj_qd, env):
        is_cuda = rj_q.is_cuda
        T, bs, nq = rj_q.shape
        n_tab = env.joint_X_p.shape[0] // env.nb
        assert n_tab in (1, bs), "per-env joint_X_p has %d blocks for %d envs" % (n_tab, bs)
        q = _f32c(rj_q, env.device).reshape(T * bs, nq)
        if rj_qd is None:
            qd = torch.zeros(T * bs, nq - 1, device=env.device, dtype=torch.float32)
        else:
            qd = _f32c(rj_qd, env.device).reshape(T * bs, nq - 1)
        bq, bqd = env.fk(q, qd)
        ctx.env, ctx.shape, ctx.is_cuda = env, (T, bs, nq), is_cuda
        ctx.save_for_backward(q, qd)
        body_q = bq.view(T, bs, env.nb, 7).permute(1, 0, 2, 3).contiguous()
        body_qd = bqd.view(T, bs, env.nb, 6).permute(1, 0, 2, 3).contiguous()
        body_q_numpy = LazyFrames(body_q[0])
        if not is_cuda:
            body_q, body_qd = body_q.cpu(), body_qd.cpu()
        return body_q, body_qd, body_q_numpy

    @staticmethod
    def backward(ctx, adj_body_q, adj_body_qd, _):
        env = ctx.env
        T, bs, nq = ctx.shape
        q, qd = ctx.saved_tensors
        z = lambda a, c: torch.zeros(bs, T, env.nb, c, device=env.device) if a is None else a
        aq = _f32c(z(adj_body_q, 7), env.device).permute(1, 0, 2, 3).contiguous().view(T * bs, env.nb, 7)
        aqd = _f32c(z(adj_body_qd, 6), env.device).permute(1, 0, 2, 3).contiguous().view(T * bs, env.nb, 6)
        gq, gqd = env.fk_backward(q, qd, aq, aqd)
        # reference post-processing (dp_model.py:1109-1110,1122-1123): NaN -> 0 (done by the kernel at the store),
        # upper clamp at +1 only
        gq = gq.clamp_(max=1.0).view(T, bs, nq)
        gqd = gqd.clamp_(max=1.0).view(T, bs, nq - 1)
        if not ctx.is_cuda:
            gq, gqd = gq.cpu(), gqd.cpu()
        return (gq if ctx.needs_input_grad[0] else None, gqd if ctx.needs_input_grad[1] else None, None)


def _rollout_inputs(self, q_init, qd_init, torques, res_f, refs, target_ke, target_kd, body_inv_mass, body_inertia,
                    body_inv_inertia):
    """Shared preamble of the rollout Functions: geometry from the caller object, contiguous fp32 device tensors,
    shared-parameter detection."""
    env = self.env
    dev = env.device
    bs = int(self.num_envs)
    nsteps = len(self.steps_idx)
    f2s = list(self.frame2step)
    stride = (f2s[1] - f2s[0]) if len(f2s) > 1 else nsteps
    assert all(s == i * stride for i, s in enumerate(f2s)), "frame2step must be evenly strided"
    a = dict(q_init=_f32c(q_init, dev), qd_init=_f32c(qd_init, dev), torques=_f32c(torques, dev),
             res_f=_f32c(res_f, dev), refs=_f32c(refs, dev), ke=_f32c(target_ke, dev), kd=_f32c(target_kd, dev),
             inv_m=_f32c(body_inv_mass, dev), I=_f32c(body_inertia, dev), inv_I=_f32c(body_inv_inertia, dev))
    assert a["q_init"].numel() == bs * env.nq and a["qd_init"].numel() == bs * env.nqd
    assert a["refs"].numel() == nsteps * bs * env.nqd
    n_tab = env.joint_X_p.shape[0] // env.nb
    assert n_tab in (1, bs), "per-env joint_X_p has %d blocks for %d envs" % (n_tab, bs)
    # Extension of the reference signature: the five parameter tensors may be given UN-replicated
    # ([nqd], [nqd], [nb], [nb,3,3], [nb,3,3]); the kernels then read one shared copy and the returned
    # gradients are summed over envs (= the backward of dp_model.py:723-725's repeat()).
    per_env = dict(ke=bs * env.nqd, kd=bs * env.nqd, inv_m=bs * env.nb, I=bs * env.nb * 9, inv_I=bs * env.nb * 9)
    shared = all(a[k].numel() * bs == n for k, n in per_env.items()) and bs > 1
    if not shared:
        for k, n in per_env.items():
            if a[k].numel() * bs == n and bs > 1:  # mixed: replicate the shared ones
                a[k] = a[k].reshape(1, -1).expand(bs, -1).reshape(-1).contiguous()
            assert a[k].numel() == n, "bad size for %s" % k
    return env, dev, bs, nsteps, stride, a, shared


class ForwardWarp(torch.autograd.Function):
    """``ForwardWarp.apply(q_init, qd_init, torques, res_f, refs, target_ke, target_kd, body_mass, body_inv_mass,
    body_inertia, body_inv_inertia, self) -> (wp_pos[F,bs*nb,7], wp_vel[F,bs*nb,6])`` -- dp_model.py:1145-1400.

    ``self`` is the caller object of the reference (``phys_model``): it must expose ``env`` (a SimEnv),
    ``num_envs``, ``steps_idx``, ``frame2step`` and ``dt``; ``grfs``, ``jafs`` and ``sim_trajs`` are written onto
    it exactly like the reference does (dp_model.py:1207-1208,1233-1234,1237,1244).  ``torques`` / ``res_f`` may be
    None (exact zeros -- the reference multiplies them by 0, dp_model.py:529,536)."""

    @staticmethod
    def forward(ctx, q_init, qd_init, torques, res_f, refs, target_ke, target_kd, body_mass, body_inv_mass,
                body_inertia, body_inv_inertia, self):
        env, dev, bs, nsteps, stride, a, shared = _rollout_inputs(self, q_init, qd_init, torques, res_f, refs, target_ke,
                                                                  target_kd, body_inv_mass, body_inertia, body_inv_inertia)
        want_forces = bool(getattr(self, "record_forces", True))
        pos, vel, grf, jaf, ws = env.rollout_forward(bs, nsteps, stride, float(self.dt), a["q_init"], a["qd_init"],
                                                     a["torques"], a["res_f"], a["refs"], a["ke"], a["kd"],
                                                     a["inv_m"], a["I"], a["inv_I"], want_forces=want_forces,
                                                     shared_params=shared)
        ctx.shared = shared
        ctx.snap = env._snapshot()
        F = pos.shape[0]
        self.grfs = [grf[i] for i in range(F)] if want_forces else []
        self.jafs = [jaf[i] for i in range(F)] if want_forces else []
        self.sim_trajs = LazyFrames(pos[:, : env.nb])
        ctx.args, ctx.ws, ctx.env, ctx.dims = a, ws, env, (bs, nsteps, stride, float(self.dt))
        ctx.shapes = dict(q_init=q_init.shape, qd_init=qd_init.shape, refs=refs.shape,
                          torques=None if torques is None else torques.shape,
                          res_f=None if res_f is None else res_f.shape, ke=target_ke.shape, kd=target_kd.shape,
                          mass=body_mass.shape, inv_m=body_inv_mass.shape, I=body_inertia.shape,
                          inv_I=body_inv_inertia.shape)
        return pos, vel

    @staticmethod
    def backward(ctx, adj_body_qs, adj_body_qd):
        env, a = ctx.env, ctx.args
        bs, nsteps, stride, dt = ctx.dims
        if env._snapshot() != ctx.snap:
            # the adjoint re-reads the model (checkpoint layout, attach gains, gravity, ground, joint_X_p): it must still
            # be the one the forward pass saw, otherwise the checkpoint rows would be misread / a different model
            # differentiated without any error
            raise _lib.PprError("SimEnv was modified between ForwardWarp.forward and .backward "
                                "(checkpoint policy, latency layout, attach gains, gravity, ground or joint_X_p)")
        fn = env.rollout_backward_shared if ctx.shared else env.rollout_backward
        g = fn(bs, nsteps, stride, dt, a["q_init"], a["qd_init"], a["torques"], a["res_f"], a["refs"], a["ke"], a["kd"],
               a["inv_m"], a["I"], a["inv_I"], _f32c(adj_body_qs, env.device), _f32c(adj_body_qd, env.device), ctx.ws)
        need, sh = ctx.needs_input_grad, ctx.shapes

        def pick(i, t, shape):
            if not need[i] or t is None:
                return None
            # remove_nan (dp_utils.py:43-57, clip=False) is applied by the adjoint kernel at every gradient store (nan0)
            return t.reshape(shape)    # (un-replicated parameters: already summed over envs on the device)
        body_mass_grad = torch.zeros(sh["mass"], device=env.device) if need[7] else None  # K5 never reads m
        return (pick(0, g["q_init"], sh["q_init"]), pick(1, g["qd_init"], sh["qd_init"]),
                pick(2, g["torques"], sh["torques"]), pick(3, g["res_f"], sh["res_f"]),
                pick(4, g["refs"], sh["refs"]), pick(5, g["target_ke"], sh["ke"]), pick(6, g["target_kd"], sh["kd"]),
                body_mass_grad, pick(8, g["body_inv_mass"], sh["inv_m"]), pick(9, g["body_inertia"], sh["I"]),
                pick(10, g["body_inv_inertia"], sh["inv_I"]), None)


class Se3Loss(torch.autograd.Function):
    """``Se3Loss.apply(pred[...,7|6], gt[...,7|6], rot_ratio) -> loss[...]`` -- the reference's ``se3_loss``
    (dp_utils.py:113-138; poses xyz + quaternion xyzw, or twists xyz + axis-angle) as ONE kernel forward and one
    backward instead of the ~330-400 elementwise torch kernels of the composed expression."""

    @staticmethod
    def forward(ctx, pred, gt, rot_ratio=0.1):
        assert pred.is_cuda and pred.shape == gt.shape and pred.shape[-1] in (6, 7)
        dim = pred.shape[-1]
        p = _f32c(pred, pred.device)
        g = _f32c(gt, pred.device)
        n = p.numel() // dim
        loss = torch.empty(pred.shape[:-1], device=pred.device, dtype=torch.float32)
        with torch.cuda.device(pred.device):
            _lib.check(_lib.lib().ppr_se3_loss_forward(n, dim, _ptr(p), _ptr(g), C.c_float(rot_ratio), _ptr(loss),
                                                       _stream()), "ppr_se3_loss_forward")
        ctx.save_for_backward(p, g)
        ctx.rot_ratio, ctx.dim = float(rot_ratio), dim
        return loss

    @staticmethod
    def backward(ctx, adj_loss):
        p, g = ctx.saved_tensors
        dim = ctx.dim
        n = p.numel() // dim
        a = _f32c(adj_loss, p.device)
        ap = torch.empty_like(p)
        ag = torch.empty_like(g) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib().ppr_se3_loss_backward(n, dim, _ptr(p), _ptr(g), C.c_float(ctx.rot_ratio), _ptr(a),
                                                        _ptr(ap), _ptr(ag), _stream()), "ppr_se3_loss_backward")
        return (ap if ctx.needs_input_grad[0] else None), ag, None


class FrameCompose(torch.autograd.Function):
    """``FrameCompose.apply(global_q[7], q[...,7], delta[...,6]) -> (target[...,7], queried[...,7])`` --
    ``rotate_frame(global_q, q)`` followed by ``compose_delta(target, delta)`` of the batch-input producer
    (dp_utils.py:60-72,21-30 inside get_batch_input, dp_model.py:611-662) as one kernel forward and one backward
    (SURVEY.md 8f rank 2).  Gradients flow to ``global_q`` and ``delta``; ``q`` (mocap data) gets none."""

    @staticmethod
    def forward(ctx, global_q, q, delta):
        assert q.is_cuda and q.shape[-1] == 7 and delta.shape[-1] == 6 and q.shape[:-1] == delta.shape[:-1]
        dev = q.device
        g, qq, d = _f32c(global_q, dev), _f32c(q, dev), _f32c(delta, dev)
        n = qq.numel() // 7
        target, queried = torch.empty_like(qq), torch.empty_like(qq)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().ppr_frame_compose_forward(n, _ptr(g), _ptr(qq), _ptr(d), _ptr(target), _ptr(queried),
                                                            _stream()), "ppr_frame_compose_forward")
        ctx.save_for_backward(g, qq, d)
        return target, queried

    @staticmethod
    def backward(ctx, adj_target, adj_queried):
        g, qq, d = ctx.saved_tensors
        n = qq.numel() // 7
        z = lambda a: torch.zeros_like(qq) if a is None else _f32c(a, qq.device)
        at, aq = z(adj_target), z(adj_queried)
        adj_g = torch.empty(n, 7, device=qq.device, dtype=torch.float32)
        adj_d = torch.empty_like(d)
        with torch.cuda.device(qq.device):
            _lib.check(_lib.lib().ppr_frame_compose_backward(n, _ptr(g), _ptr(qq), _ptr(d), _ptr(at), _ptr(aq),
                                                             _ptr(adj_g), _ptr(adj_d), _stream()),
                       "ppr_frame_compose_backward")
        return (adj_g.sum(0) if ctx.needs_input_grad[0] else None), None, (adj_d if ctx.needs_input_grad[2] else None)


class RefsFromFrames(torch.autograd.Function):
    """``RefsFromFrames.apply(frames[F, n], stride, T) -> refs[T, n]``: per-substep control references linearly
    interpolated from per-frame values on the device -- what ``get_mocap_data``'s scipy ``interp1d`` does on the host for
    every substep of every window (dp_model.py:421-427,605-609).  A caller ships F x n floats instead of T x n."""

    @staticmethod
    def forward(ctx, frames, stride, T):
        assert frames.is_cuda and frames.dim() == 2 and frames.shape[0] >= (T - 1) // stride + 1
        f = _f32c(frames, frames.device)
        refs = torch.empty(T, f.shape[1], device=f.device, dtype=torch.float32)
        with torch.cuda.device(f.device):
            _lib.check(_lib.lib().ppr_refs_from_frames(T, stride, f.shape[0], f.shape[1], _ptr(f), _ptr(refs), _stream()),
                       "ppr_refs_from_frames")
        ctx.dims = (T, stride, f.shape[0], f.shape[1])
        return refs

    @staticmethod
    def backward(ctx, adj_refs):
        T, stride, F, n = ctx.dims
        a = _f32c(adj_refs, adj_refs.device)
        adj = torch.empty(F, n, device=a.device, dtype=torch.float32)
        with torch.cuda.device(a.device):
            _lib.check(_lib.lib().ppr_refs_from_frames_backward(T, stride, F, n, _ptr(a), _ptr(adj), _stream()),
                       "ppr_refs_from_frames_backward")
        return adj, None, None


def _addr(t):
    return None if t is None else t.data_ptr()


class ForwardWarpLoss(torch.autograd.Function):
    """``ForwardWarpLoss.apply(q_init, qd_init, torques, res_f, refs, target_ke, target_kd, body_mass, body_inv_mass,
    body_inertia, body_inv_inertia, target_pos[F, bs*nb, 7], rot_ratio, self) -> (loss_pos[F, bs*nb], wp_pos, wp_vel)``

    ForwardWarp with the imitation objective's pose loss FUSED into the rollout:
    ``loss_pos = se3_loss(wp_pos, target_pos, rot_ratio)`` per body and frame (dp_model.py:777, dp_utils.py:113-138) is
    evaluated by the forward kernel while the pose is in registers, and the adjoint kernel seeds itself from
    ``d objective / d loss_pos`` -- no ``adj_pos`` tensor, no separate loss kernels; the gradient w.r.t. ``target_pos``
    (the targets depend on ``global_q``) is written by the adjoint kernel as well.  The caller applies its own mean /
    masking / clipping (``reduce_loss``) to ``loss_pos``.  ``wp_pos`` / ``wp_vel`` are returned for the terms that use the
    simulated states detached (dp_model.py:794,800) and for ``sim_trajs``; they carry NO gradient here."""

    @staticmethod
    def forward(ctx, q_init, qd_init, torques, res_f, refs, target_ke, target_kd, body_mass, body_inv_mass, body_inertia,
                body_inv_inertia, target_pos, rot_ratio, self):
        env, dev, bs, nsteps, stride, a, shared = _rollout_inputs(self, q_init, qd_init, torques, res_f, refs, target_ke,
                                                                  target_kd, body_inv_mass, body_inertia, body_inv_inertia)
        F = (nsteps - 1) // stride + 1
        tgt = _f32c(target_pos, dev)
        assert tgt.numel() == F * bs * env.nb * 7, "target_pos must be [F, bs*nb, 7]"
        e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        want_forces = bool(getattr(self, "record_forces", True))
        pos, vel, loss = e(F, bs * env.nb, 7), e(F, bs * env.nb, 6), e(F, bs * env.nb)
        grf = e(F, bs * env.nb, 6) if want_forces else None
        jaf = e(F, bs * env.nb, 6) if want_forces else None
        ws = e((env.workspace_bytes(bs, nsteps) + 3) // 4)
        io = RolloutIO()
        io.bs, io.nsteps, io.frame_stride, io.dt, io.shared_params = bs, nsteps, stride, float(self.dt), int(shared)
        for k, v in (("q_init", a["q_init"]), ("qd_init", a["qd_init"]), ("torques", a["torques"]), ("res_f", a["res_f"]),
                     ("refs", a["refs"]), ("target_ke", a["ke"]), ("target_kd", a["kd"]), ("body_inv_mass", a["inv_m"]),
                     ("body_inertia", a["I"]), ("body_inv_inertia", a["inv_I"]), ("out_pos", pos), ("out_vel", vel),
                     ("out_grf", grf), ("out_jaf", jaf), ("workspace", ws), ("target_pos", tgt), ("loss_pos", loss)):
            setattr(io, k, _addr(v))
        io.workspace_bytes, io.rot_ratio = ws.numel() * 4, float(rot_ratio)
        with torch.cuda.device(dev):
            _lib.check(env._lib.ppr_rollout_forward_ex(env._h, C.byref(io), _stream()), "ppr_rollout_forward_ex")
        self.grfs = [grf[i] for i in range(F)] if want_forces else []
        self.jafs = [jaf[i] for i in range(F)] if want_forces else []
        self.sim_trajs = LazyFrames(pos[:, : env.nb])
        ctx.env, ctx.a, ctx.keep, ctx.shared, ctx.snap = env, a, (pos, vel, ws, tgt), shared, env._snapshot()
        ctx.dims = (bs, nsteps, stride, float(self.dt), float(rot_ratio))
        ctx.shapes = dict(q_init=q_init.shape, qd_init=qd_init.shape, refs=refs.shape,
                          torques=None if torques is None else torques.shape,
                          res_f=None if res_f is None else res_f.shape, ke=target_ke.shape, kd=target_kd.shape,
                          mass=body_mass.shape, inv_m=body_inv_mass.shape, I=body_inertia.shape, inv_I=body_inv_inertia.shape,
                          target=target_pos.shape)
        ctx.mark_non_differentiable(pos, vel)
        return loss, pos, vel

    @staticmethod
    def backward(ctx, adj_loss, _adj_pos, _adj_vel):
        env, a = ctx.env, ctx.a
        bs, nsteps, stride, dt, rot_ratio = ctx.dims
        if env._snapshot() != ctx.snap:
            raise _lib.PprError("SimEnv was modified between ForwardWarpLoss.forward and .backward")
        pos, vel, ws, tgt = ctx.keep
        dev, nb, nq, nqd = env.device, env.nb, env.nq, env.nqd
        e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        al = _f32c(adj_loss, dev)
        g = dict(q_init=e(bs * nq), qd_init=e(bs * nqd), refs=e(nsteps, bs * nqd),
                 torques=e(nsteps, bs * nqd) if a["torques"] is not None else None,
                 res_f=e(nsteps, bs * nb, 6) if a["res_f"] is not None else None)
        io = RolloutIO()
        io.bs, io.nsteps, io.frame_stride, io.dt, io.shared_params = bs, nsteps, stride, dt, int(ctx.shared)
        keep = []
        if ctx.shared:
            P = 2 * nqd + 19 * nb
            packed = e(P)
            scratch = e((int(env._lib.ppr_rollout_reduce_scratch_bytes(env._h, bs)) + 3) // 4)
            io.adj_shared, io.reduce_scratch, io.reduce_scratch_bytes = packed.data_ptr(), scratch.data_ptr(), scratch.numel() * 4
            g.update(target_ke=packed[:nqd], target_kd=packed[nqd:2 * nqd], body_inv_mass=packed[2 * nqd:2 * nqd + nb],
                     body_inertia=packed[2 * nqd + nb:2 * nqd + 10 * nb].view(nb, 3, 3),
                     body_inv_inertia=packed[2 * nqd + 10 * nb:].view(nb, 3, 3))
            keep += [packed, scratch]
        else:
            g.update(target_ke=e(bs * nqd), target_kd=e(bs * nqd), body_inv_mass=e(bs * nb), body_inertia=e(bs * nb, 3, 3),
                     body_inv_inertia=e(bs * nb, 3, 3))
            for k, f in (("target_ke", "adj_target_ke"), ("target_kd", "adj_target_kd"), ("body_inv_mass", "adj_body_inv_mass"),
                         ("body_inertia", "adj_body_inertia"), ("body_inv_inertia", "adj_body_inv_inertia")):
                setattr(io, f, g[k].data_ptr())
        for k, v in (("q_init", a["q_init"]), ("qd_init", a["qd_init"]), ("torques", a["torques"]), ("res_f", a["res_f"]),
                     ("refs", a["refs"]), ("target_ke", a["ke"]), ("target_kd", a["kd"]), ("body_inv_mass", a["inv_m"]),
                     ("body_inertia", a["I"]), ("body_inv_inertia", a["inv_I"]), ("out_pos", pos), ("out_vel", vel),
                     ("workspace", ws), ("target_pos", tgt), ("adj_loss_pos", al), ("adj_q_init", g["q_init"]),
                     ("adj_qd_init", g["qd_init"]), ("adj_torques", g["torques"]), ("adj_res_f", g["res_f"]),
                     ("adj_refs", g["refs"])):
            setattr(io, k, _addr(v))
        io.workspace_bytes, io.rot_ratio = ws.numel() * 4, rot_ratio
        g_tgt = None
        if ctx.needs_input_grad[11]:     # the targets depend on global_q (and on the root-pose net through FK)
            g_tgt = torch.empty_like(tgt)
            io.adj_target_pos = g_tgt.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(env._lib.ppr_rollout_backward_ex(env._h, C.byref(io), _stream()), "ppr_rollout_backward_ex")
        need, sh = ctx.needs_input_grad, ctx.shapes
        pick = lambda i, t, shape: None if (not need[i] or t is None) else t.reshape(shape)
        body_mass_grad = torch.zeros(sh["mass"], device=dev) if need[7] else None  # K5 never reads m
        return (pick(0, g["q_init"], sh["q_init"]), pick(1, g["qd_init"], sh["qd_init"]),
                pick(2, g["torques"], sh["torques"]), pick(3, g["res_f"], sh["res_f"]), pick(4, g["refs"], sh["refs"]),
                pick(5, g["target_ke"], sh["ke"]), pick(6, g["target_kd"], sh["kd"]), body_mass_grad,
                pick(8, g["body_inv_mass"], sh["inv_m"]), pick(9, g["body_inertia"], sh["I"]),
                pick(10, g["body_inv_inertia"], sh["inv_I"]), None if g_tgt is None else g_tgt.reshape(sh["target"]), None,
                None)
