"""Consumes ``tests/golden/warp_<robot>.npz`` -- outputs of the UNMODIFIED reference running on Warp, written by
tools/dump_warp_reference.py on a machine where ``warp_lang==0.7.2`` is installed.  No such file can be produced in this
project's containers (no Warp, no network), so these tests are skipped until one is committed; with the files present
they pin the oracle (and the model compiler) against Warp itself and turn "parity unpinned" into a checked claim.

The reading side is exercised regardless: ``test_dump_reader_on_an_oracle_made_stand_in`` writes a file of the same
layout from the oracle and runs it through the same checks."""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import oracle_rollout_grads, rel_err

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DUMPS = sorted(glob.glob(os.path.join(GOLDEN, "warp_*.npz")))
KEYS10 = ["q_init", "qd_init", "torques", "res_f", "refs", "target_ke", "target_kd", "body_inv_mass", "body_inertia",
          "body_inv_inertia"]


def load_dump(path):
    from ppr_diffphys_b200.model import RobotModel, _ARRAY_FIELDS
    z = np.load(path, allow_pickle=False)
    rm = RobotModel(name=str(z["model_name"]), joint_attach_ke=float(z["model_joint_attach_ke"]),
                    joint_attach_kd=float(z["model_joint_attach_kd"]),
                    body_names=[str(s) for s in z["model_body_names"]], **{k: z["model_" + k] for k in _ARRAY_FIELDS})
    bs = z["in_q_init"].size // rm.nq
    T = z["in_refs"].shape[0]
    d = dict(q_init=z["in_q_init"].reshape(bs, rm.nq), qd_init=z["in_qd_init"].reshape(bs, rm.nqd),
             torques=z["in_torques"].reshape(T, bs, rm.nqd), res_f=z["in_res_f"].reshape(T, bs, rm.nb, 6),
             refs=z["in_refs"].reshape(T, bs, rm.nqd), target_ke=z["in_target_ke"].reshape(bs, rm.nqd),
             target_kd=z["in_target_kd"].reshape(bs, rm.nqd), body_inv_mass=z["in_body_inv_mass"].reshape(bs, rm.nb),
             body_inertia=z["in_body_inertia"].reshape(bs, rm.nb, 3, 3),
             body_inv_inertia=z["in_body_inv_inertia"].reshape(bs, rm.nb, 3, 3))
    return z, rm, {k: torch.from_numpy(np.asarray(v)).double() for k, v in d.items()}, bs, T


def check_oracle_against_dump(path, pos_tol=1e-4, grad_tol=None):
    """float64 oracle on the dumped model + inputs vs what Warp (float32) produced."""
    from helpers import fp32_noise_floor
    z, rm, d, bs, T = load_dump(path)
    stride, F = int(z["stride"]), int(z["nframes"])
    adj_pos, adj_vel = torch.from_numpy(z["adj_pos"]).double(), torch.from_numpy(z["adj_vel"]).double()
    pos, vel, g = oracle_rollout_grads(rm, d, stride, F, adj_pos=adj_pos, adj_vel=adj_vel, dt=float(z["dt"]))
    perr = float((pos.reshape(F, -1, 7) - torch.from_numpy(z["pos"]).double()).abs().max())
    assert perr <= pos_tol, perr
    floor, _ = fp32_noise_floor(rm, d, stride, F, adj_pos=adj_pos, adj_vel=adj_vel, dt=float(z["dt"]))
    for k in KEYS10:
        tol = grad_tol if grad_tol is not None else max(1e-3, 2.0 * floor[k])
        e = rel_err(torch.from_numpy(z["grad_" + k]), g[k])
        assert e <= tol, (k, e, tol)
    assert np.abs(z["grad_body_mass"]).max() == 0.0        # K5 reads m but never uses it (integrator_euler.py:43)
    return rm


@pytest.mark.skipif(not DUMPS, reason="no tests/golden/warp_*.npz (needs a machine with warp_lang==0.7.2: "
                                      "tools/dump_warp_reference.py)")
@pytest.mark.parametrize("path", DUMPS or ["-"])
def test_oracle_matches_warp_dump(path):
    rm = check_oracle_against_dump(path)
    # model compiler fidelity: this repo's compiled asset vs Warp's own builder output
    from ppr_diffphys_b200 import load_robot
    mine = load_robot(rm.name)
    assert (mine.nb, mine.nq, mine.nqd) == (rm.nb, rm.nq, rm.nqd)
    assert list(mine.joint_type) == list(rm.joint_type) and list(mine.joint_parent) == list(rm.joint_parent)
    assert np.allclose(mine.joint_X_p, rm.joint_X_p, atol=1e-6) and np.allclose(mine.joint_axis, rm.joint_axis, atol=1e-6)
    assert np.allclose(mine.body_mass, rm.body_mass, rtol=1e-4) and np.allclose(mine.body_com, rm.body_com, atol=1e-5)
    assert np.allclose(mine.norm_body_inertia, rm.norm_body_inertia, rtol=1e-3, atol=1e-7)
    assert mine.nc == rm.nc, (mine.nc, rm.nc)


@pytest.mark.gpu
@pytest.mark.skipif(not DUMPS, reason="no tests/golden/warp_*.npz")
@pytest.mark.parametrize("path", DUMPS or ["-"])
def test_cuda_matches_warp_dump(path):
    from helpers import fp32_noise_floor
    from ppr_diffphys_b200 import SimEnv
    from test_gpu_parity import flat_args, run_cuda
    z, rm, d, bs, T = load_dump(path)
    stride, F = int(z["stride"]), int(z["nframes"])
    env = SimEnv(rm)
    a, _, _ = flat_args(d, torch.device("cuda:0"))
    pos, vel, _ = run_cuda(env, a, bs, T, stride)
    assert float((pos.cpu().double() - torch.from_numpy(z["pos"]).double()).abs().max()) <= 1e-4
    dev = pos.device
    torch.autograd.backward([pos, vel], [torch.from_numpy(z["adj_pos"]).to(dev), torch.from_numpy(z["adj_vel"]).to(dev)])
    floor, _ = fp32_noise_floor(rm, d, stride, F, adj_pos=torch.from_numpy(z["adj_pos"]).double(),
                                adj_vel=torch.from_numpy(z["adj_vel"]).double())
    for k in KEYS10:
        e = rel_err(a[k].grad, torch.from_numpy(z["grad_" + k]))
        assert e <= max(1e-3, 4.0 * floor[k]), (k, e)       # two independent fp32 evaluations: 2 x floor each


def test_dump_reader_on_an_oracle_made_stand_in(tmp_path):
    """A file with the dump's layout, filled by the float64 oracle instead of Warp, goes through the same reader and
    checks (keeps the consuming code alive while no real dump exists)."""
    from helpers import make_inputs, settle_height
    from ppr_diffphys_b200 import load_robot
    from ppr_diffphys_b200.model import _ARRAY_FIELDS
    rm = load_robot("human")
    stride, F, bs = 8, 3, 2
    T = stride * (F - 1) + 1
    rm, d = make_inputs(rm, bs=bs, T=T, seed=21, lin_vel=0.5, res_f_std=0.05, torque_std=0.05, ang=0.25)
    d = settle_height(rm, d, 0.003)
    d = {k: v.float().double() for k, v in d.items()}
    g = torch.Generator().manual_seed(11)
    wp = torch.randn(F, bs * rm.nb, 7, generator=g, dtype=torch.float64)
    wv = 0.1 * torch.randn(F, bs * rm.nb, 6, generator=g, dtype=torch.float64)
    pos, vel, grads = oracle_rollout_grads(rm, d, stride, F, adj_pos=wp, adj_vel=wv)
    out = {"model_" + k: getattr(rm, k) for k in _ARRAY_FIELDS}
    out.update(model_name=rm.name, model_joint_attach_ke=rm.joint_attach_ke, model_joint_attach_kd=rm.joint_attach_kd,
               model_body_names=np.array(rm.body_names))
    flat = dict(q_init=d["q_init"].reshape(-1), qd_init=d["qd_init"].reshape(-1), torques=d["torques"].reshape(T, -1),
                res_f=d["res_f"].reshape(T, -1, 6), refs=d["refs"].reshape(T, -1), target_ke=d["target_ke"].reshape(-1),
                target_kd=d["target_kd"].reshape(-1), body_mass=d["body_mass"].reshape(-1),
                body_inv_mass=d["body_inv_mass"].reshape(-1), body_inertia=d["body_inertia"].reshape(-1, 3, 3),
                body_inv_inertia=d["body_inv_inertia"].reshape(-1, 3, 3))
    out.update({"in_" + k: v.numpy().astype(np.float32) for k, v in flat.items()})
    out.update(pos=pos.reshape(F, -1, 7).numpy().astype(np.float32), vel=vel.reshape(F, -1, 6).numpy().astype(np.float32),
               adj_pos=wp.numpy().astype(np.float32), adj_vel=wv.numpy().astype(np.float32))
    shapes = dict(q_init=(-1,), qd_init=(-1,), torques=(T, -1), res_f=(T, -1, 6), refs=(T, -1), target_ke=(-1,),
                  target_kd=(-1,), body_inv_mass=(-1,), body_inertia=(-1, 3, 3), body_inv_inertia=(-1, 3, 3))
    out.update({"grad_" + k: grads[k].reshape(shapes[k]).numpy().astype(np.float32) for k in KEYS10})
    out.update(grad_body_mass=np.zeros(bs * rm.nb, np.float32), dt=5e-4, stride=stride, nframes=F, robot="human")
    path = os.path.join(str(tmp_path), "warp_standin.npz")
    np.savez_compressed(path, **out)
    check_oracle_against_dump(path, grad_tol=1e-5)
