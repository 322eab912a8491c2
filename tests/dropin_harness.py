"""Drop-in harness: runs the REFERENCE's own ``phys_model`` class (diffphys/dp_model.py:56-1011) on top of this
repo's operators.

At test time (never committed) the reference's ``dp_model.py`` and the pure-torch helper modules it imports are copied
from a reference checkout into a temporary package, the edits of INTEGRATION.md section 1 are applied as exact string
replacements (each anchor must be found exactly once -- a reference that drifted fails loudly), and the third-party
modules that are not installable here are shimmed:

  dqtorch      three quaternion conversions (geom_utils.py:17-200, dp_utils.py:129-133)          -> torch, below
  trimesh      imported by dataloader.py / lab4d_utils.py, never called on this path              -> empty module
  urdfpy       reached only through diffphys/robot.py and diffphys/urdf_utils.py (mesh posing for the visualiser and
               ``get_foot_height``, dp_model.py:574-579)                                          -> both modules are
               replaced by a stand-in that poses the compiled robot's collision vertices (for laikago the contact
               points ARE the collision-mesh vertices, so the lowest vertex is the same quantity)
  warp         gone after the edits (no import left)

Backends: ``cuda`` = the product (SimEnv on libppr_b200.so); ``cpu-double`` = ``CpuSimEnv`` below, a SimEnv look-alike
on the oracle's C++ CPU port, so that the very same autograd Functions of ppr_diffphys_b200/ops.py (shape handling,
shared-parameter detection, gradient selection, side channels) run under the reference class on a box without a GPU.
Test infrastructure only.
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("PPR_REFERENCE_ROOT", "/root/reference")
COPIED = ["dp_model.py", "dp_utils.py", "geom_utils.py", "torch_utils.py", "dataloader.py", "lab4d_utils.py"]


def reference_available():
    return all(os.path.exists(os.path.join(REF_ROOT, "diffphys", f)) for f in COPIED)


# ------------------------------------------------------------------------------------------------ the edits
def _replace_once(src, old, new, what):
    n = src.count(old)
    assert n == 1, "INTEGRATION.md edit '%s': anchor found %d times in the reference" % (what, n)
    return src.replace(old, new)


def _cut(src, start, end, new, what):
    i = src.index(start) if src.count(start) == 1 else -1
    assert i >= 0, "INTEGRATION.md edit '%s': start anchor not unique" % what
    j = src.index(end, i)
    return src[:i] + new + src[j:]


def apply_integration_edits(src):
    """INTEGRATION.md section 1 on the text of diffphys/dp_model.py."""
    # imports: Warp / parse_urdf / integrator  ->  this package
    src = _replace_once(src, "from warp.sim.articulation import eval_fk\n"
                             "from diffphys.import_urdf import parse_urdf\n"
                             "from diffphys.integrator_euler import SemiImplicitIntegrator\n",
                        "from ppr_diffphys_b200 import ForwardKinematics, ForwardWarp, SimEnv, compile_robot\n", "imports")
    src = _replace_once(src, "import warp as wp\n\nwp.init()\n", "", "warp init")
    # edit 1a: model build in __init__ (ModelBuilder + parse_urdf + mass / inertia post-processing + integrator, :125-208)
    src = _cut(src, "        # env\n        self.articulation_builder = wp.sim.ModelBuilder()\n",
               "        self.target_ke = nn.Parameter(\n",
               "        # env: ONE compiled copy of the static model (INTEGRATION.md edit 1)\n"
               "        self.robot_model = compile_robot(opts[\"urdf_template\"], \"%s/data/urdf_templates\" % data_dir)\n"
               "        self.articulation_builder = self.robot_model.as_builder_view()\n"
               "        self.n_dof = self.robot_model.nq - 7\n"
               "        self.n_links = self.robot_model.nb\n\n", "model build")
    # edit 1b: the one-env model of init_global_q (:246-250)
    src = _replace_once(src, "        builder = wp.sim.ModelBuilder()\n"
                             "        for i in range(self.num_envs):\n"
                             "            builder.add_rigid_articulation(self.articulation_builder)\n"
                             "        self.env = builder.finalize(self.device)\n\n        # mocap data\n",
                        "        self.env = SimEnv(self.robot_model)\n\n        # mocap data\n", "init_global_q env")
    # edit 1c: reinit_envs (:384-401): no replicated builder, no State list, no collide()
    src = _cut(src, "            builder = wp.sim.ModelBuilder()\n            for i in range(self.num_envs):\n",
               "            setattr(self, env_name, self.env)\n",
               "            self.env = SimEnv(self.robot_model)\n"
               "            self.env.ground = True\n\n"
               "            self.env.joint_attach_ke = self.joint_attach_ke\n"
               "            self.env.joint_attach_kd = self.joint_attach_kd\n"
               "            self.state_steps = None   # the checkpoint workspace lives on the autograd ctx\n\n", "reinit_envs env")
    # edit 2: delete the local ForwardKinematics / wp_add / ForwardWarp (:1022-1400); convert_ppr_warp (:1014) stays
    i = src.index("class ForwardKinematics(torch.autograd.Function):")
    src = src[:i]
    live = [l for l in src.splitlines() if "wp." in l and not l.lstrip().startswith("#")]
    assert not live, "a Warp call survived the edits: %r" % live[:3]
    return src


# ------------------------------------------------------------------------------------------------ shims
DQTORCH_SHIM = '''
"""Stand-in for the three dqtorch functions the reference calls (real-first quaternions)."""
import torch


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def _sqrt_pos(x):
    return torch.where(x > 0, torch.sqrt(torch.clamp(x, min=1e-30)), torch.zeros_like(x))


def matrix_to_quaternion(m):
    b = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(b + (9,)), -1)
    q_abs = _sqrt_pos(torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22,
                                   1 - m00 - m11 + m22], -1))
    cand = torch.stack([torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
                        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
                        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
                        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1)], -2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    idx = q_abs.argmax(-1)
    return torch.gather(cand, -2, idx[..., None, None].expand(b + (1, 4))).squeeze(-2)


def axis_angle_to_quaternion(v):
    ang = v.norm(p=2, dim=-1, keepdim=True)
    half = 0.5 * ang
    small = ang.abs() < 1e-6
    s = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    return torch.cat([torch.cos(half), v * s], -1)
'''

ROBOT_SHIM = '''
"""Stand-in for diffphys/robot.py + diffphys/urdf_utils.py (urdfpy / trimesh mesh posing: visualiser and foot height)."""
import numpy as np
import torch


class _Urdf:
    pass


class URDFRobot:
    def __init__(self, urdf_path):
        self.urdf = _Urdf()          # no ``kp_links`` attribute for laikago (robot.py:55-96 sets it for human / quad only)
        self.urdf_path = urdf_path
'''

URDF_UTILS_SHIM = '''
import torch
from diffphys.geom_utils import se3_vec2mat

_MODEL = {}


def bind_robot_model(rm):
    _MODEL["rm"] = rm


def articulate_robot_rbrt_batch(robot, rbrt):
    """urdf_utils.py:154-200: collision-mesh vertices of every body posed by rbrt[..., nb, 7] -> (verts[..., V, 3], faces)."""
    rm = _MODEL["rm"]
    cb = torch.as_tensor(rm.contact_body, dtype=torch.long, device=rbrt.device)
    cp = torch.as_tensor(rm.contact_point, dtype=rbrt.dtype, device=rbrt.device)
    T = se3_vec2mat(rbrt[..., cb, :])                       # ..., V, 4, 4
    verts = (T[..., :3, :3] @ cp[..., None])[..., 0] + T[..., :3, 3]
    return verts, None


def articulate_robot_rbrt(*a, **k):
    raise NotImplementedError("visualiser path (out of scope of the drop-in test)")


articulate_robot = articulate_robot_rbrt
'''


def build_patched_reference(tmp_dir, ref_root=REF_ROOT):
    """tmp_dir/refpkg/diffphys/{patched dp_model, copied helpers, shims}; returns the sys.path entry to prepend."""
    root = os.path.join(str(tmp_dir), "refpkg")
    pkg = os.path.join(root, "diffphys")
    os.makedirs(pkg, exist_ok=True)
    for f in COPIED:
        shutil.copy(os.path.join(ref_root, "diffphys", f), os.path.join(pkg, f))
    src = open(os.path.join(pkg, "dp_model.py")).read()
    open(os.path.join(pkg, "dp_model.py"), "w").write(apply_integration_edits(src))
    open(os.path.join(pkg, "__init__.py"), "w").write("")
    open(os.path.join(pkg, "robot.py"), "w").write(ROBOT_SHIM)
    open(os.path.join(pkg, "urdf_utils.py"), "w").write(URDF_UTILS_SHIM)
    open(os.path.join(root, "dqtorch.py"), "w").write(DQTORCH_SHIM)
    open(os.path.join(root, "trimesh.py"), "w").write("# stand-in: imported by the reference, never called on this path\n")
    if not os.path.exists(os.path.join(root, "data")):
        os.symlink(os.path.join(ref_root, "data"), os.path.join(root, "data"))   # URDFs + mocap clips, read in place
    return root


def import_patched(root):
    for name in [m for m in sys.modules if m == "diffphys" or m.startswith("diffphys.") or m in ("dqtorch", "trimesh")]:
        del sys.modules[name]
    sys.path.insert(0, root)
    try:
        return importlib.import_module("diffphys.dp_model"), importlib.import_module("diffphys.dataloader"), \
            importlib.import_module("diffphys.urdf_utils")
    finally:
        sys.path.remove(root)


def default_opts(tmp_dir, seqname="mi-pace", urdf_template="laikago"):
    """run.sh:12 + the flag defaults of main.py:15-41."""
    return dict(seqname=seqname, logname="0", logroot=os.path.join(str(tmp_dir), "logdir"), urdf_template=urdf_template,
                phys_learning_rate=1e-4, num_rounds=5, warmup_iters=0, iters_per_round=20, ratio_phys_cycle=1.0,
                noise_std=2e-3, traj_wt=0.01, pos_state_wt=0.01, vel_state_wt=1e-4, pos_distill_wt=0.0,
                reg_torque_wt=0.0, reg_res_f_wt=0.0, reg_foot_wt=0.0, reg_root_wt=0.0, accu_steps=1)


# ------------------------------------------------------------------------------------------------ CPU double of SimEnv
class CpuSimEnv:
    """Same surface as ppr_diffphys_b200.ops.SimEnv (what the autograd Functions and the reference class touch), on the
    oracle's C++ CPU port."""

    def __init__(self, robot, device=None):
        from oracle.cpu_port import CpuRollout
        from ppr_diffphys_b200.model import load_robot
        self.model = load_robot(robot) if isinstance(robot, str) else robot
        self.device = torch.device("cpu")
        self.cpu = CpuRollout(self.model)
        self.nb, self.nq, self.nqd = self.model.nb, self.model.nq, self.model.nqd
        self.body_count_per_env = self.nb
        self.ground = True
        self.joint_attach_ke = float(self.model.joint_attach_ke)
        self.joint_attach_kd = float(self.model.joint_attach_kd)
        self.joint_X_p = torch.as_tensor(self.model.joint_X_p).clone()
        self.body_com = torch.as_tensor(self.model.body_com)
        self.checkpoint_every = 1

    def _snapshot(self):
        return (self.joint_attach_ke, self.joint_attach_kd, self.ground)

    def fk(self, q, qd):
        return self.cpu.fk(q, qd)

    def fk_backward(self, q, qd, abq, abqd):
        return self.cpu.fk_backward(q, qd, abq, abqd)

    def rollout_forward(self, bs, nsteps, stride, dt, q_init, qd_init, torques, res_f, refs, ke, kd, inv_m, I, inv_I,
                        want_forces=True, workspace=None, shared_params=False):
        nb, nq, nqd = self.nb, self.nq, self.nqd
        F = (nsteps - 1) // stride + 1
        rep = (lambda t, *s: t.reshape(1, *s).expand(bs, *s).contiguous()) if shared_params else (lambda t, *s: t.reshape(bs, *s))
        d = dict(q_init=q_init.reshape(bs, nq), qd_init=qd_init.reshape(bs, nqd),
                 torques=None if torques is None else torques.reshape(nsteps, bs, nqd),
                 res_f=None if res_f is None else res_f.reshape(nsteps, bs, nb, 6), refs=refs.reshape(nsteps, bs, nqd),
                 target_ke=rep(ke, nqd), target_kd=rep(kd, nqd), body_inv_mass=rep(inv_m, nb),
                 body_inertia=rep(I, nb, 3, 3), body_inv_inertia=rep(inv_I, nb, 3, 3))
        out = self.cpu.forward(d, dt, stride, F, want_forces=want_forces)
        pos, vel = out[0].reshape(F, bs * nb, 7), out[1].reshape(F, bs * nb, 6)
        grf = out[2].reshape(F, bs * nb, 6) if want_forces else None
        jaf = out[3].reshape(F, bs * nb, 6) if want_forces else None
        return pos, vel, grf, jaf, torch.zeros(1)

    def rollout_backward(self, bs, nsteps, stride, dt, q_init, qd_init, torques, res_f, refs, ke, kd, inv_m, I, inv_I,
                         adj_pos, adj_vel, workspace, shared_params=False):
        F = (nsteps - 1) // stride + 1
        g = self.cpu.backward(adj_pos.reshape(F, bs, self.nb, 7), adj_vel.reshape(F, bs, self.nb, 6))
        flat = lambda t: None if t is None else t.reshape(-1) if t.dim() <= 2 else t
        return dict(q_init=g["q_init"].reshape(-1), qd_init=g["qd_init"].reshape(-1),
                    torques=None if g["torques"] is None else g["torques"].reshape(nsteps, -1),
                    res_f=None if g["res_f"] is None else g["res_f"].reshape(nsteps, bs * self.nb, 6),
                    refs=g["refs"].reshape(nsteps, -1), target_ke=g["target_ke"].reshape(-1),
                    target_kd=g["target_kd"].reshape(-1), body_inv_mass=g["body_inv_mass"].reshape(-1),
                    body_inertia=g["body_inertia"].reshape(-1, 3, 3), body_inv_inertia=g["body_inv_inertia"].reshape(-1, 3, 3))


def run_reference_loop(backend, tmp_dir, iters, num_envs=10, frames_per_wdw=24, monkeypatch=None):
    """Patched reference class: construct, reinit_envs, then `iters` x (forward, backward, update) like main.py:64-105.
    Returns (model, list of loss dicts as floats, list of grad dicts)."""
    import ppr_diffphys_b200 as pkg
    root = build_patched_reference(tmp_dir)
    if backend == "cpu-double":
        monkeypatch.setattr(pkg, "SimEnv", CpuSimEnv)
        monkeypatch.setattr(torch.cuda, "LongTensor", torch.LongTensor, raising=False)   # dp_model.py:361 hard-codes CUDA
        device = "cpu"
    else:
        device = "cuda"
    dp_model, dataloader, urdf_utils = import_patched(root)
    opts = default_opts(tmp_dir)
    os.makedirs(os.path.join(opts["logroot"], "%s-%s" % (opts["seqname"], opts["logname"])), exist_ok=True)
    cwd = os.getcwd()
    os.chdir(REF_ROOT)     # the reference opens ./data/motion_sequences/... relative to its root (dataloader.py:13)
    try:
        loader = dataloader.DataLoader(opts)
    finally:
        os.chdir(cwd)
    np.random.seed(0)
    torch.manual_seed(0)
    # bind the compiled robot for the mesh-posing stand-in before the constructor calls get_foot_height
    orig_compile = pkg.compile_robot

    def compile_and_bind(name, urdf_root):
        rm = orig_compile(name, urdf_root)
        urdf_utils.bind_robot_model(rm)
        return rm
    dp_model.compile_robot = compile_and_bind
    model = dp_model.phys_model(opts, loader, device=device)
    if device == "cuda":
        model.cuda()
    model.train()
    losses, grads = [], []
    for it in range(iters):
        model.progress = it / (opts["num_rounds"] * opts["iters_per_round"])
        if it % opts["iters_per_round"] == 0:
            model.save_checkpoint(it)
            model.reinit_envs(num_envs, frames_per_wdw=frames_per_wdw, is_eval=False)
        loss_dict = model.forward()
        model.backward(loss_dict["total_loss"])
        grad_dict = model.update()
        losses.append({k: float(v) for k, v in loss_dict.items()})
        grads.append({k: float(v) for k, v in grad_dict.items()})
    return model, losses, grads
