#!/usr/bin/env python
"""Route to a PINNED oracle: dump what the unmodified reference computes with Warp, on a machine where it runs.

Runs only where ``warp_lang==0.7.2`` (requirements.txt:12), ``urdfpy`` and ``dqtorch`` import and a checkout of
gengshan-y/ppr-diffphys is available -- none of which is true in this project's build container or on its GPU boxes
(no network).  For each robot it writes ``tests/golden/warp_<robot>.npz`` holding

  * the static arrays of the Warp ``Model`` the reference builds (``parse_urdf`` + the post-processing of
    dp_model.py:123-205 + ``reinit_envs``, :384-401) under the field names of ``ppr_diffphys_b200.RobotModel`` -- the
    answer to SURVEY section 8c's "builder rules are recalled, not verified";
  * seeded inputs (``tests/helpers.make_inputs``, float32-rounded) and the reference's own ``ForwardWarp.apply`` outputs
    on them: body poses / twists at the frame steps and the eleven gradients of a fixed linear loss
    (dp_model.py:1147-1400, ``wp.Tape``), plus ``ForwardKinematics`` on the initial state.

``tests/test_warp_dump.py`` consumes those files when they exist (skipped otherwise): oracle vs Warp on the CPU, CUDA
vs Warp with ``-m gpu``.  Commit the .npz files to turn "parity unpinned" into a pinned oracle.

    python tools/dump_warp_reference.py --reference /path/to/ppr-diffphys [--device cuda|cpu] [--robots laikago human quad]
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

KEYS = ["q_init", "qd_init", "torques", "res_f", "refs", "target_ke", "target_kd", "body_mass", "body_inv_mass",
        "body_inertia", "body_inv_inertia"]          # the eleven tensor arguments of ForwardWarp.apply, in order
FRAMES_PER_WDW, BS = 3, 4                          # T = 33 * 2 + 1 = 67 substeps with the shipped clips (dt 5e-4)


def warp_array(a):
    return np.array(a.numpy() if hasattr(a, "numpy") else a)


def dump_model(model, nb):
    """Warp Model (num_envs articulations) -> RobotModel fields of articulation 0."""
    env, b = model.env, model.articulation_builder
    nq, nqd = len(b.joint_q), len(b.joint_qd)
    nc = int(env.contact_count) // int(model.num_envs)
    f32, i32 = (lambda x: np.ascontiguousarray(np.asarray(x, np.float32))), (lambda x: np.ascontiguousarray(np.asarray(x, np.int32)))
    shape_body = np.asarray(b.shape_body)
    return dict(
        name=model.opts["urdf_template"], joint_type=i32(warp_array(env.joint_type)[:nb]),
        joint_parent=i32(warp_array(env.joint_parent)[:nb]), joint_X_p=f32(warp_array(env.joint_X_p)[:nb]),
        joint_X_c=f32(warp_array(env.joint_X_c)[:nb]), joint_axis=f32(warp_array(env.joint_axis)[:nb]),
        joint_q_start=i32(warp_array(env.joint_q_start)[:nb]), joint_qd_start=i32(warp_array(env.joint_qd_start)[:nb]),
        joint_limit_lower=f32(warp_array(env.joint_limit_lower)[:nqd]), joint_limit_upper=f32(warp_array(env.joint_limit_upper)[:nqd]),
        joint_limit_ke=f32(warp_array(env.joint_limit_ke)[:nqd]), joint_limit_kd=f32(warp_array(env.joint_limit_kd)[:nqd]),
        joint_target_ke=f32(model.target_ke.detach().cpu()), joint_target_kd=f32(model.target_kd.detach().cpu()),
        body_com=f32(warp_array(env.body_com)[:nb]), body_mass=f32(model.body_mass.detach().cpu()),
        norm_body_inertia=f32(model.norm_body_inertia.detach().cpu()),
        contact_body=i32(warp_array(env.contact_body0)[:nc]), contact_point=f32(warp_array(env.contact_point0)[:nc]),
        contact_dist=f32(warp_array(env.contact_dist)[:nc]),
        # Warp indexes shape_materials by SHAPE (contact_material = shape index); keep its table for articulation 0
        contact_material=i32(warp_array(env.contact_material)[:nc]),
        shape_materials=f32(warp_array(env.shape_materials)[: len(shape_body)]),
        gravity=f32(warp_array(env.gravity) if hasattr(env.gravity, "__len__") else env.gravity),
        joint_q_rest=f32(b.joint_q), joint_attach_ke=float(env.joint_attach_ke), joint_attach_kd=float(env.joint_attach_kd),
        body_names=np.array([str(n) for n in getattr(b, "body_name", [""] * nb)][:nb]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("PPR_REFERENCE_ROOT", "/root/reference"))
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--robots", nargs="+", default=["laikago", "human", "quad"])
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    try:
        import warp as wp
    except ImportError:
        sys.exit("warp_lang is not installed here: this tool only runs where the reference itself runs "
                 "(pip install warp-lang==0.7.2 urdfpy==0.0.22 trimesh==3.9.43 + dqtorch, see the reference README)")
    sys.path.insert(0, args.reference)
    os.chdir(args.reference)                      # the reference opens ./data/... relative to its root
    from diffphys import dp_model as ref
    from diffphys.dataloader import DataLoader
    from helpers import make_inputs, settle_height
    from dropin_harness import default_opts
    if args.device == "cpu":                      # dp_model.py hard-codes CUDA in a few places (:361, :1034-1035)
        torch.cuda.LongTensor = torch.LongTensor
        torch.Tensor.cuda = lambda self, *a, **k: self
    # the constructor also builds MLPs / optimiser and needs laikago-shaped mocap for its FK probe: skip those parts,
    # everything that builds the simulator model (:76-222) still runs unmodified
    for name in ("add_nn_modules", "init_global_q", "add_optimizer"):
        setattr(ref.phys_model, name, lambda self, *a, **k: None)
    os.makedirs(args.out, exist_ok=True)
    for robot in args.robots:
        opts = default_opts("/tmp", seqname="mi-pace", urdf_template=robot)
        model = ref.phys_model(opts, DataLoader(opts), device=args.device)
        model.reinit_envs(BS, frames_per_wdw=FRAMES_PER_WDW, is_eval=False, overwrite=True)
        nb = model.n_links
        arrays = dump_model(model, nb)
        T, stride = len(model.steps_idx), model.steps_per_fr_interval
        # ---- seeded inputs on the DUMPED model (so that oracle / CUDA later see identical parameters)
        from ppr_diffphys_b200.model import RobotModel, _ARRAY_FIELDS
        rm = RobotModel(name=robot, joint_attach_ke=arrays["joint_attach_ke"], joint_attach_kd=arrays["joint_attach_kd"],
                        body_names=list(arrays["body_names"]), **{k: arrays[k] for k in _ARRAY_FIELDS})
        rm, d = make_inputs(rm, bs=BS, T=T, seed=21, lin_vel=0.5, res_f_std=0.05, torque_std=0.05, ang=0.25)
        ja = d["q_init"][:, 7:]
        d["q_init"][:, 7:] = torch.where(ja.abs() < 0.05, 0.05 * torch.sign(ja) + (ja == 0) * 0.05, ja)
        d = settle_height(rm, d, penetration=0.003)
        d = {k: v.float() for k, v in d.items()}
        dev = torch.device(args.device)
        flat = dict(q_init=d["q_init"].reshape(-1), qd_init=d["qd_init"].reshape(-1), torques=d["torques"].reshape(T, -1),
                    res_f=d["res_f"].reshape(T, -1, 6), refs=d["refs"].reshape(T, -1), target_ke=d["target_ke"].reshape(-1),
                    target_kd=d["target_kd"].reshape(-1), body_mass=d["body_mass"].reshape(-1),
                    body_inv_mass=d["body_inv_mass"].reshape(-1), body_inertia=d["body_inertia"].reshape(-1, 3, 3),
                    body_inv_inertia=d["body_inv_inertia"].reshape(-1, 3, 3))
        a = {k: v.to(dev).contiguous().requires_grad_(True) for k, v in flat.items()}
        pos, vel = ref.ForwardWarp.apply(*[a[k] for k in KEYS], model)
        g = torch.Generator().manual_seed(11)
        wp_ = torch.randn(pos.shape, generator=g).to(dev)
        wv_ = (0.1 * torch.randn(vel.shape, generator=g)).to(dev)
        ((pos * wp_).sum() + (vel * wv_).sum()).backward()
        # FK of the initial state through the reference's own Function (layout [T=1, bs, nq])
        bq, bqd, _ = ref.ForwardKinematics.apply(d["q_init"][None].to(dev), d["qd_init"][None].to(dev), model.env)
        out = {("model_" + k): v for k, v in arrays.items()}
        out.update({"in_" + k: flat[k].numpy() for k in KEYS})
        out.update(pos=pos.detach().cpu().numpy(), vel=vel.detach().cpu().numpy(), adj_pos=wp_.cpu().numpy(),
                   adj_vel=wv_.cpu().numpy(), fk_body_q=bq.detach().cpu().numpy(), fk_body_qd=bqd.detach().cpu().numpy(),
                   grfs=np.stack([x.detach().cpu().numpy() for x in model.grfs]),
                   jafs=np.stack([x.detach().cpu().numpy() for x in model.jafs]))
        out.update({"grad_" + k: (np.zeros(flat[k].shape, np.float32) if a[k].grad is None else a[k].grad.cpu().numpy())
                    for k in KEYS})
        out.update(dt=float(model.dt), stride=int(stride), nframes=int(FRAMES_PER_WDW), robot=robot,
                   warp_version=str(getattr(wp.config, "version", "unknown")), device=args.device)
        path = os.path.join(args.out, "warp_%s.npz" % robot)
        np.savez_compressed(path, **out)
        print("wrote", path, "nb", nb, "contacts/env", len(arrays["contact_body"]), "T", T)


if __name__ == "__main__":
    main()
