"""The reference's own ``phys_model`` (diffphys/dp_model.py:56-1011), patched with the edits of INTEGRATION.md at test
time, running reinit_envs / forward / backward / update (main.py:64-105) over this repo's operators.

* ``cpu-double``: no GPU needed -- SimEnv is the CPU-port look-alike of tests/dropin_harness.py, the autograd Functions
  are the product's.  Runs wherever a reference checkout exists (this build container).
* ``cuda``: the product end to end.  Needs a GPU AND a reference checkout (``PPR_REFERENCE_ROOT``); the GPU boxes of this
  project have no reference tree, so it is skipped there.
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dropin_harness as H  # noqa: E402

needs_ref = pytest.mark.skipif(not H.reference_available(), reason="no reference checkout at %s" % H.REF_ROOT)

LOSS_KEYS = {"loss_traj", "loss_pos_state", "loss_vel_state", "loss_reg_torque", "loss_reg_res_f", "loss_reg_foot",
             "total_loss"}                                                       # dp_model.py:775-836
STATE_KEYS = {"target_ke", "target_kd", "body_mass", "norm_body_inertia", "global_q"}   # dp_model.py:210-222,263-267
MLPS = ("root_pose_mlp", "joint_angle_mlp", "vel_mlp", "torque_mlp", "residual_f_mlp")  # dp_model.py:292-315


@needs_ref
def test_integration_edits_apply_to_the_reference():
    src = open(os.path.join(H.REF_ROOT, "diffphys", "dp_model.py")).read()
    out = H.apply_integration_edits(src)
    compile(out, "dp_model_patched.py", "exec")
    assert "from ppr_diffphys_b200 import ForwardKinematics, ForwardWarp, SimEnv, compile_robot" in out
    assert "import warp" not in out
    # everything of the caller class is still the reference's: forward, losses, optimiser, checkpoint queue
    for name in ("def forward(self, frame_start=None)", "def get_batch_input(self, steps_fr)", "def check_grad(self",
                 "def save_checkpoint(self", "def convert_ppr_warp(tensor)", "ForwardWarp.apply(", "ForwardKinematics.apply("):
        assert name in out, name
    assert len(out.splitlines()) > 850        # only the builder block and the three Warp classes are gone


def _check_run(model, losses, grads, iters):
    assert len(losses) == iters
    assert set(losses[0]) == LOSS_KEYS
    sd = model.state_dict()
    assert STATE_KEYS <= set(sd), STATE_KEYS - set(sd)
    for m in MLPS:
        assert any(k.startswith(m + ".") for k in sd), m
    assert all(torch.isfinite(torch.tensor(list(l.values()))).all() for l in losses)
    assert any(k.startswith("grad/") for k in grads[-1])          # check_grad ran and the optimiser stepped
    # side channels the reference's query() consumes (dp_model.py:855-874)
    F = model.frames_per_wdw
    assert len(model.sim_trajs) == F and model.sim_trajs[0].shape == (model.n_links, 7)
    assert len(model.grfs) == F and tuple(model.grfs[0].shape) == (model.num_envs * model.n_links, 6)
    assert len(model.target_trajs) == F and len(model.pid_ref) == F


@needs_ref
def test_reference_phys_model_trains_on_the_operator_interface(tmp_path, monkeypatch):
    iters = 20
    model, losses, grads = H.run_reference_loop("cpu-double", tmp_path, iters, num_envs=10, frames_per_wdw=24,
                                                monkeypatch=monkeypatch)
    _check_run(model, losses, grads, iters)
    assert model.n_links == 13 and model.n_dof == 12 and len(model.steps_idx) == 33 * 23 + 1     # 760 substeps
    # mi-pace converges (README: iteration 0 vs 100): mean trajectory loss of the last 5 iterations < first 5
    first = sum(l["loss_traj"] for l in losses[:5]) / 5
    last = sum(l["loss_traj"] for l in losses[-5:]) / 5
    print("loss_traj first5 %.5f last5 %.5f" % (first, last))
    assert last < first
    # the evaluation pass of main.py:77-79: ONE env over the whole clip (39 frames = 1 255 substeps), no noise
    model.eval()
    model.reinit_envs(1, frames_per_wdw=model.total_frames, is_eval=True)
    with torch.no_grad():
        ev = model.forward()
    assert len(model.steps_idx) == 33 * 38 + 1 and len(model.sim_trajs) == model.total_frames == 39
    assert set(ev) == LOSS_KEYS and all(torch.isfinite(v) for v in ev.values())
    # back to the (cached) training envs, like the next line of main.py
    model.train()
    model.reinit_envs(10, frames_per_wdw=24, is_eval=False)
    assert model.num_envs == 10 and len(model.steps_idx) == 760


@needs_ref
@pytest.mark.gpu
def test_reference_phys_model_trains_on_cuda(tmp_path, monkeypatch):
    iters = 20
    model, losses, grads = H.run_reference_loop("cuda", tmp_path, iters, num_envs=10, frames_per_wdw=24,
                                                monkeypatch=monkeypatch)
    _check_run(model, losses, grads, iters)
    first = sum(l["loss_traj"] for l in losses[:5]) / 5
    last = sum(l["loss_traj"] for l in losses[-5:]) / 5
    assert last < first
