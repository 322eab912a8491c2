"""Host-side (CPU tensor) checks of the caller loop's loss / geometry helpers (ppr_diffphys_b200/imitation.py), the
torch code that surrounds the two drop-in ops exactly as dp_model.py / dp_utils.py do in the reference."""
import math

import numpy as np
import torch

from ppr_diffphys_b200 import imitation as im


def test_reduce_loss_masked_mean_equals_reference_semantics():
    """dp_utils.py:93-110: mean over the entries > 0, over all entries when none is -- the capture-safe masked mean
    must return the same value and the same gradient."""
    g = torch.Generator().manual_seed(0)
    for case in range(4):
        x = torch.rand(7, 5, generator=g, dtype=torch.float64)
        if case == 1:
            x[x < 0.5] = 0.0
        if case == 2:
            x.zero_()
        if case == 3:
            x[0, 0] = 0.0
        x.requires_grad_(True)
        ref = x[x > 0].mean() if bool((x > 0).any()) else x.mean()
        out = im.reduce_loss(x)
        assert torch.allclose(out, ref, atol=1e-15)
        if case != 2:
            g1, = torch.autograd.grad(out, x)
            g2, = torch.autograd.grad(ref, x)
            assert torch.allclose(g1, g2, atol=1e-15)


def test_quaternion_helpers_xyzw():
    g = torch.Generator().manual_seed(1)
    q = torch.randn(20, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    p = torch.randn(20, 4, generator=g, dtype=torch.float64)
    p = p / p.norm(dim=-1, keepdim=True)
    R = im.quat_to_matrix
    assert torch.allclose(R(im.quat_mul(q, p)), R(q) @ R(p), atol=1e-12)          # homomorphism
    assert torch.allclose(R(q) @ R(q).transpose(-1, -2), torch.eye(3, dtype=torch.float64).expand(20, 3, 3), atol=1e-12)
    # axis-angle -> quaternion: rotation about z by 90 degrees maps x to y; small-angle branch is continuous
    qz = im.axis_angle_to_quat(torch.tensor([0.0, 0.0, math.pi / 2], dtype=torch.float64))
    assert torch.allclose(R(qz) @ torch.tensor([1.0, 0, 0], dtype=torch.float64),
                          torch.tensor([0.0, 1, 0], dtype=torch.float64), atol=1e-12)
    a = im.axis_angle_to_quat(torch.tensor([0.0, 0.0, 0.99e-6], dtype=torch.float64))
    b = im.axis_angle_to_quat(torch.tensor([0.0, 0.0, 1.01e-6], dtype=torch.float64))
    assert torch.allclose(a, b, atol=1e-7) and abs(float(a.norm()) - 1) < 1e-12


def test_se3_loss_poses_and_twists():
    """dp_utils.py:113-138: squared translation error + 0.1 * geodesic rotation angle; NaN rows contribute 0."""
    ident = torch.tensor([0.0, 0, 0, 0, 0, 0, 1.0], dtype=torch.float64)
    ang = 0.3
    other = torch.tensor([1.0, 2.0, -1.0, 0, math.sin(ang / 2), 0, math.cos(ang / 2)], dtype=torch.float64)
    l = im.se3_loss(other[None], ident[None])
    assert abs(float(l) - (6.0 + 0.1 * ang)) < 1e-9
    assert float(im.se3_loss(ident[None], ident[None])) <= 0.1 * math.acos(1 - 1e-4) + 1e-12   # rot_angle's eps clamp
    tw_a = torch.tensor([0.0, 0, 0, 0.2, 0, 0], dtype=torch.float64)
    tw_b = torch.tensor([0.5, 0, 0, 0.0, 0, 0], dtype=torch.float64)
    assert abs(float(im.se3_loss(tw_a[None], tw_b[None])) - (0.25 + 0.1 * 0.2)) < 1e-9
    bad = other.clone()
    bad[0] = float("nan")
    assert float(im.se3_loss(bad[None], ident[None])) == 0.0


def test_rotate_frame_and_compose_delta():
    g = torch.Generator().manual_seed(2)
    q = torch.randn(6, 7, generator=g, dtype=torch.float64)
    q[:, 3:] = q[:, 3:] / q[:, 3:].norm(dim=-1, keepdim=True)
    ident = torch.tensor([0.0, 0, 0, 0, 0, 0, 1.0], dtype=torch.float64)
    assert torch.allclose(im.rotate_frame(ident, q), q, atol=1e-12)
    assert torch.allclose(im.compose_delta(q, torch.zeros(6, 6, dtype=torch.float64)), q, atol=1e-12)
    # T_global @ T: translation part
    gq = torch.tensor([0.1, 0.2, 0.3, 0, 0, math.sin(0.25), math.cos(0.25)], dtype=torch.float64)
    out = im.rotate_frame(gq, q)
    R = im.quat_to_matrix(gq[3:])
    assert torch.allclose(out[:, :3], q[:, :3] @ R.T + gq[:3], atol=1e-12)
    assert torch.allclose(im.quat_to_matrix(out[:, 3:]), R @ im.quat_to_matrix(q[:, 3:]), atol=1e-12)
    # velocities rotate, do not translate
    qd = torch.randn(6, 6, generator=g, dtype=torch.float64)
    assert torch.allclose(im.rotate_frame_vel(gq, qd)[:, :3], qd[:, :3] @ R.T, atol=1e-12)


def test_motion_clip_assets_and_parse():
    frames, dur = im.load_motion("mi-pace")
    assert frames.shape[1] >= 49 and abs(dur - 0.01667) < 1e-4          # 39 frames, 1/60 s
    m = im.parse_amp(torch.as_tensor(frames))
    assert m["pos"].shape[-1] == 3 and m["orn"].shape[-1] == 4 and m["jang"].shape[-1] == 12
    assert np.allclose(np.linalg.norm(m["orn"].numpy(), axis=-1), 1.0, atol=1e-4)
