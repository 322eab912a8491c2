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
    l = im.se3_loss_torch(other[None], ident[None])
    assert abs(float(l) - (6.0 + 0.1 * ang)) < 1e-9
    assert float(im.se3_loss_torch(ident[None], ident[None])) <= 0.1 * math.acos(1 - 1e-4) + 1e-12   # rot_angle's eps clamp
    tw_a = torch.tensor([0.0, 0, 0, 0.2, 0, 0], dtype=torch.float64)
    tw_b = torch.tensor([0.5, 0, 0, 0.0, 0, 0], dtype=torch.float64)
    assert abs(float(im.se3_loss_torch(tw_a[None], tw_b[None])) - (0.25 + 0.1 * 0.2)) < 1e-9
    bad = other.clone()
    bad[0] = float("nan")
    assert float(im.se3_loss_torch(bad[None], ident[None])) == 0.0


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


def test_se3_loss_adjoint_f64_equals_autograd():
    """ppr_loss.h (the scalar-templated loss + hand-written adjoint the CUDA kernels instantiate in float32) compiled
    in float64 inside the CPU port == torch autograd of the composed definition, including small-angle twists, a
    clamped (identical) pair and a NaN row."""
    import ctypes as C
    from oracle import cpu_port
    lib = cpu_port.lib()
    g = torch.Generator().manual_seed(0)
    for dim in (7, 6):
        n = 300
        pred = torch.randn(n, dim, generator=g, dtype=torch.float64)
        gt = torch.randn(n, dim, generator=g, dtype=torch.float64)
        if dim == 6:
            pred[:5, 3:] = 0
            gt[:5, 3:] *= 1e-9
            pred[5:10, 3:] *= 1e-8
        pred[20] = gt[20]
        pred[21, 0] = float("nan")
        pred.requires_grad_(True)
        gt.requires_grad_(True)
        w = torch.rand(n, generator=g, dtype=torch.float64)
        l = im.se3_loss_torch(pred, gt)
        (l * w).sum().backward()
        loss, ap, ag = np.zeros(n), np.zeros((n, dim)), np.zeros((n, dim))
        P, G, W = pred.detach().numpy().copy(), gt.detach().numpy().copy(), w.numpy().copy()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        lib.ppr_cpu_se3_loss_f64(C.c_int64(n), dim, vp(P), vp(G), C.c_double(0.1), vp(loss), vp(W), vp(ap), vp(ag))
        ok = np.ones(n, bool)
        ok[21] = False                      # NaN row: loss 0 and zero gradient here, NaN gradient in torch
        if dim == 6:
            ok[:10] = False                 # |v| -> 0: torch's norm() has no gradient at 0; checked separately
        assert np.abs(loss - l.detach().numpy())[ok].max() < 1e-12
        assert np.abs(ap - pred.grad.numpy())[ok].max() < 1e-11 and np.abs(ag - gt.grad.numpy())[ok].max() < 1e-11
        assert loss[21] == 0 and not ap[21].any() and not ap[20, 3:].any()
        assert np.isfinite(ap).all() and np.isfinite(ag).all()
        if dim == 6:
            assert np.abs(loss - l.detach().numpy())[:10].max() < 1e-12
            assert np.abs(ap[5:10] - pred.grad.numpy()[5:10]).max() < 1e-9   # tiny but non-zero angles


def test_se3_loss_float32_accuracy_beats_trace_form():
    """The float32 instantiation of ppr_loss.h (relative-quaternion angle, what the CUDA kernel runs) against the
    float64 truth: at least as accurate as the float32 trace / acos form of the composed definition, by 1-2 orders of
    magnitude at the small angles of the training loop.  Pairs within rounding of rot_angle's clamp boundary are
    excluded (the gradient is discontinuous there)."""
    import ctypes as C
    from oracle import cpu_port
    lib = cpu_port.lib()
    g = torch.Generator().manual_seed(3)
    for dim in (7, 6):
        for noise in (0.3, 0.03):
            shape = (10, 24, 13, dim)
            pred = torch.randn(shape, generator=g)
            gt = pred + noise * torch.randn(shape, generator=g)
            if dim == 7:
                gt[..., 3:] = gt[..., 3:] / gt[..., 3:].norm(dim=-1, keepdim=True)
            w = torch.rand(shape[:-1], generator=g)
            n = pred.numel() // dim

            def torch_grads(dt):
                p = pred.to(dt).clone().requires_grad_(True)
                l = im.se3_loss_torch(p, gt.to(dt))
                (l * w.to(dt)).sum().backward()
                return l.detach().double().numpy().reshape(n), p.grad.double().numpy().reshape(n, dim)

            l64, g64 = torch_grads(torch.float64)
            l32, g32 = torch_grads(torch.float32)
            P, G, W = pred.numpy().reshape(n, dim).copy(), gt.numpy().reshape(n, dim).copy(), w.numpy().reshape(n).copy()
            loss, ap = np.zeros(n, np.float32), np.zeros((n, dim), np.float32)
            vp = lambda a: a.ctypes.data_as(C.c_void_p)
            lib.ppr_cpu_se3_loss_f32(C.c_int64(n), dim, vp(P), vp(G), C.c_float(0.1), vp(loss), vp(W), vp(ap), None)
            rp, rg = pred.double()[..., 3:], gt.double()[..., 3:]
            if dim == 6:
                rp, rg = im.axis_angle_to_quat(rp), im.axis_angle_to_quat(rg)
            cos = ((im.quat_to_matrix(rp) * im.quat_to_matrix(rg)).sum((-1, -2)) - 1) / 2
            ok = (~((cos - (1 - 1e-4)).abs() < 5e-6)).reshape(n).numpy()
            scale = np.abs(g64).max()
            err_fused = np.abs(ap.astype(np.float64) - g64)[ok].max() / scale
            err_torch = np.abs(g32 - g64)[ok].max() / scale
            assert err_fused < 1e-5 and err_fused <= err_torch, (dim, noise, err_fused, err_torch)
            assert np.abs(loss.astype(np.float64) - l64)[ok].max() <= max(np.abs(l32 - l64)[ok].max(), 5e-7)


def test_frame_compose_adjoint_f64_equals_autograd():
    """ppr_frame.h (rotate_frame + compose_delta of the batch-input producer and their hand-written adjoint, what the
    CUDA kernels instantiate in float32) in float64 == the composed torch functions + autograd, including zero and
    tiny axis-angle deltas and an un-normalised global quaternion."""
    import ctypes as C
    from oracle import cpu_port
    lib = cpu_port.lib()
    g = torch.Generator().manual_seed(0)
    n = 200
    gq = torch.tensor([0.1, 0.2, -0.3, 0.05, -0.1, 0.2, 1.1], dtype=torch.float64, requires_grad=True)
    q = torch.randn(n, 7, generator=g, dtype=torch.float64)
    q[:, 3:] /= q[:, 3:].norm(dim=-1, keepdim=True)
    d = torch.randn(n, 6, generator=g, dtype=torch.float64) * 0.3
    d[:3, 3:] = 0
    d[3:6, 3:] *= 1e-8
    d.requires_grad_(True)
    t = im.rotate_frame(gq, q)
    u = im.compose_delta(t, d)
    wt = torch.randn(n, 7, generator=g, dtype=torch.float64)
    wu = torch.randn(n, 7, generator=g, dtype=torch.float64)
    ((t * wt).sum() + (u * wu).sum()).backward()
    T, U, ag, ad = np.zeros((n, 7)), np.zeros((n, 7)), np.zeros((n, 7)), np.zeros((n, 6))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    GQ, Q, D = gq.detach().numpy().copy(), q.numpy().copy(), d.detach().numpy().copy()
    WT, WU = wt.numpy().copy(), wu.numpy().copy()
    lib.ppr_cpu_frame_compose_f64(C.c_int64(n), vp(GQ), vp(Q), vp(D), vp(T), vp(U), vp(WT), vp(WU), vp(ag), vp(ad))
    assert np.abs(T - t.detach().numpy()).max() < 1e-13 and np.abs(U - u.detach().numpy()).max() < 1e-13
    assert np.abs(ag.sum(0) - gq.grad.numpy()).max() < 1e-11
    assert np.abs(ad - d.grad.numpy()).max() < 1e-12


def test_reduce_loss_clipping_matches_the_reference_loop():
    """Masked / graph-capturable reduce_loss(clip=True) == a literal re-statement of dp_utils.py:93-110 (threshold from
    env 0, every env cut from its first entry above it, mean over the positive entries)."""
    from ppr_diffphys_b200.imitation import reduce_loss

    def literal(loss_seq):
        loss_seq = loss_seq.clone()
        th = 0
        for i in range(len(loss_seq)):
            if th == 0:
                sub = loss_seq[i]
                th = sub[sub > 0].median() * 10
            clip_val, clip_idx = torch.max(loss_seq[i] > th, 0)
            if clip_val == 1:
                loss_seq[i, clip_idx:] = 0
        return loss_seq[loss_seq > 0].mean() if loss_seq.sum() > 0 else loss_seq.mean()

    g = torch.Generator().manual_seed(0)
    for trial in range(20):
        x = torch.rand(6, 9, generator=g) * 0.01
        if trial % 2:
            x[torch.randint(0, 6, (1,), generator=g), torch.randint(0, 9, (1,), generator=g)] = 5.0    # a diverged window
        if trial % 3 == 0:
            x[:, 0] = 0                                                                                # masked (out-of-seq) entries
        assert torch.allclose(reduce_loss(x, clip=True), literal(x), rtol=1e-6, atol=0), trial
        assert torch.allclose(reduce_loss(x), x[x > 0].mean())
    z = torch.zeros(3, 4)
    assert float(reduce_loss(z, clip=True)) == 0.0 and float(reduce_loss(z)) == 0.0


def test_median_gradient_clipping_matches_the_reference_queue():
    """median_clip_ == the per-parameter queue logic of dp_model.py:965-998 (literal re-statement with
    torch.nn.utils.clip_grad_norm_), over a sequence of gradients with outliers."""
    from ppr_diffphys_b200.imitation import median_clip_
    g = torch.Generator().manual_seed(0)
    names = ["a", "b"]
    mine_q, ref_q = {}, {}
    for it in range(40):
        grads = {n: torch.randn(7, generator=g) * (1.0 if n == "a" else 0.1) for n in names}
        if it in (15, 16, 30):
            grads["a"] = grads["a"] * 50.0                       # outliers
        ref = {}
        for n in names:                                          # the reference's loop
            p = torch.nn.Parameter(torch.zeros(7))
            p.grad = grads[n].clone()
            grad = p.grad.reshape(-1).norm(2, -1)
            q = ref_q.setdefault(n, [])
            if len(q) > 10:
                med = torch.stack(q[:-1]).median()
                if grad > 5.0 * med:
                    torch.nn.utils.clip_grad_norm_(p, med)
                else:
                    q.append(grad)
                    q.pop(0)
            else:
                q.append(grad)
            ref[n] = p.grad
        mine = {n: grads[n].clone() for n in names}
        info = median_clip_(mine.items(), mine_q)
        for n in names:
            assert torch.allclose(mine[n], ref[n], rtol=1e-5, atol=1e-7), (it, n)
            assert len(mine_q[n]) == len(ref_q[n]) and all(abs(x - float(y)) < 1e-5 * max(1.0, x) for x, y in zip(mine_q[n], ref_q[n]))
        if it in (15, 16, 30):
            assert info["a"][2] and not info["b"][2]


def test_device_median_clipping_equals_the_host_queue():
    """median_clip_device_ (tensor ops only, graph-capturable) == median_clip_ (host queue, itself checked against the
    reference's loop above) on a gradient sequence with outliers and dropped iterations."""
    from ppr_diffphys_b200.imitation import median_clip_, median_clip_device_
    g = torch.Generator().manual_seed(1)
    names = ["a", "b", "c"]
    shapes = {"a": (7,), "b": (3, 4), "c": (1,)}
    host_q = {}
    Q, cnt = torch.zeros(3, 11), torch.zeros((), dtype=torch.long)
    n_clipped = 0
    for it in range(45):
        grads = {n: torch.randn(shapes[n], generator=g) * (0.1 if n == "b" else 1.0) for n in names}
        if it in (14, 15, 31):
            grads["a"] = grads["a"] * 80.0
        if it in (20, 31):
            grads["c"] = grads["c"] * 1e3
        dropped = it in (3, 17, 33)
        host = {n: v.clone() for n, v in grads.items()}
        if not dropped:                                           # update() does not clip / enqueue a dropped iteration
            info = median_clip_(host.items(), host_q)
            n_clipped += sum(v[2] for v in info.values())
        dev = [grads[n].clone() for n in names]
        norms, clipped = median_clip_device_(dev, Q, cnt, torch.tensor(dropped))
        for n, d in zip(names, dev):
            assert torch.allclose(d, host[n], rtol=1e-5, atol=1e-8), (it, n)
        if not dropped:
            assert [bool(c) for c in clipped] == [info[n][2] for n in names], it
        else:
            assert not clipped.any()
        assert int(cnt) == len(host_q.get("a", []))
        for i, n in enumerate(names):
            assert torch.allclose(Q[i, :int(cnt)], torch.tensor(host_q.get(n, [])), rtol=1e-6, atol=0), (it, n)
    assert n_clipped >= 4 and int(cnt) == 11


def test_update_drops_rolls_back_in_place_and_steps_the_scheduler():
    """ImitationModel.update (device part + host part) on a CPU stand-in: a finite small gradient is applied; a gradient
    above the threshold or a NaN one leaves parameters, moments and step counts untouched; with a snapshot two rounds old
    the model and optimizer state roll back IN PLACE (same storage); the scheduler steps every iteration."""
    from ppr_diffphys_b200.imitation import ImitationModel

    class Toy(torch.nn.Module):
        update_device, finish_update = ImitationModel.update_device, ImitationModel.finish_update
        update, save_checkpoint = ImitationModel.update, ImitationModel.save_checkpoint

        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(4))
            self.v = torch.nn.Parameter(torch.ones(2, 2))
            self.optimizer = torch.optim.AdamW([{"params": [self.w], "lr": 1e-2}, {"params": [self.v], "lr": 1e-3}],
                                               weight_decay=1e-4, fused=True)
            self.scheduler = torch.optim.lr_scheduler.OneCycleLR(self.optimizer, [1e-2, 1e-3], 20, pct_start=0.1,
                                                                 cycle_momentum=False, anneal_strategy="linear")

    def set_grads(m, scale):
        m.w.grad = torch.full((4,), 0.1 * scale)
        m.v.grad = torch.full((2, 2), -0.2 * scale)

    m = Toy()
    ptr = (m.w.data_ptr(), m.v.data_ptr())
    set_grads(m, 1.0)
    info = m.update(keep_grads=True)
    assert not info["skipped"] and abs(info["grad_norm"] - (4 * 0.01 + 4 * 0.04) ** 0.5) < 1e-6
    assert float(m.w[0]) < 1.0 and float(m.v[0, 0]) > 1.0 and m.scheduler.last_epoch == 1
    mom_ptr = m.optimizer.state[m.w]["exp_avg"].data_ptr()
    m.save_checkpoint()                                            # snapshot A (after one step)
    w_a, step_a = m.w.detach().clone(), float(m.optimizer.state[m.w]["step"])
    exp_a = m.optimizer.state[m.w]["exp_avg"].clone()
    for bad in (1e4, float("nan")):                                # no two-rounds-old snapshot yet: drop only
        before = (m.w.detach().clone(), m.optimizer.state[m.w]["exp_avg"].clone(), float(m.optimizer.state[m.w]["step"]))
        set_grads(m, bad)
        info = m.update()
        assert info["skipped"] and m.w.grad is None
        assert torch.equal(m.w, before[0]) and torch.equal(m.optimizer.state[m.w]["exp_avg"], before[1])
        assert float(m.optimizer.state[m.w]["step"]) == before[2]
    assert m.scheduler.last_epoch == 3
    set_grads(m, 1.0)
    assert not m.update()["skipped"]
    m.save_checkpoint()                                            # snapshot B; A is now two rounds old
    set_grads(m, 0.5)
    m.update()
    assert not torch.equal(m.w, w_a)
    set_grads(m, 1e4)
    info = m.update()                                              # dropped -> roll back to A, in place
    assert info["skipped"]
    assert torch.equal(m.w, w_a) and float(m.optimizer.state[m.w]["step"]) == step_a
    assert torch.equal(m.optimizer.state[m.w]["exp_avg"], exp_a)
    assert (m.w.data_ptr(), m.v.data_ptr()) == ptr and m.optimizer.state[m.w]["exp_avg"].data_ptr() == mom_ptr
    assert m.scheduler.last_epoch == 2                             # A's epoch (1) + this iteration's scheduler step
    set_grads(m, 1.0)
    assert not m.update()["skipped"] and not torch.equal(m.w, w_a)
