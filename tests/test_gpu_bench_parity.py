"""Parity ON THE BENCHMARKED WORKLOADS: environments drawn from ``synth.make_batch`` exactly as bench.py draws them
(hovering feet that fall into contact mid-window, and the contact-heavy penetrating / sliding variant), through the same
call bench.py makes (shared un-replicated parameters, torques / res_f = None, bench.py's loss), against
``oracle.sim_oracle.rollout`` in float64.

Tolerances: body pose <= 1e-4 after the 64-substep window; every gradient within max(1e-3, 2 x the fp32 noise floor of
that gradient), the floor being the difference between the SAME oracle evaluated in float32 and in float64 on the same
inputs (helpers.fp32_noise_floor) -- i.e. never looser than twice what an independent single-precision evaluation of the
reference formulation resolves.  Achieved errors are printed (pytest -s) and asserted."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import fp32_noise_floor, oracle_rollout_grads, rel_err  # noqa: E402

pytestmark = pytest.mark.gpu
POS_TOL, GRAD_RTOL = 1e-4, 1e-3
N_PICK, N_DRAW = 16, 256

CASES = {
    # name: (bench workload, clearance override, lin_vel override)
    "laikago-65536 (+1 mm, bench headline)": ("laikago-scaling-65536x64", None, None),
    "laikago contact-heavy (-2.5 mm, sliding)": ("laikago-scaling-65536x64", -0.0025, 1.0),
    "human-65536 (contact-heavy, bench)": ("human-65536x64-contact", None, None),
    "human-4096 (contact-heavy, bench)": ("human-4096x64-contact", None, None),
    "quad-1024 (+1 mm, bench)": ("quad-1024x64", None, None),
    "quad contact-heavy (-2.5 mm, sliding)": ("quad-1024x64", -0.0025, 1.0),
}


class Caller:
    def __init__(self, env, num_envs, nsteps, stride):
        self.env, self.num_envs, self.dt = env, num_envs, 5e-4
        self.steps_idx = range(nsteps)
        self.frame2step = [i for i in range(nsteps) if i % stride == 0]
        self.record_forces = False


def bench_loss(pos, vel):
    """bench.py step(): pos [F, bs*nb, 7] / oracle [F, bs, nb, 7]"""
    return (pos[-1][..., :3] - pos[0][..., :3]).pow(2).mean() + 1e-3 * vel[-1].pow(2).mean()


@pytest.mark.parametrize("layout", ["throughput-layout", "latency-layout"])
@pytest.mark.parametrize("case", list(CASES))
def test_bench_workload_matches_oracle(case, layout, monkeypatch):
    import bench
    from ppr_diffphys_b200 import ForwardWarp, SimEnv, load_robot
    monkeypatch.setenv("PPR_LATENCY_ENVS", "0" if layout == "throughput-layout" else "1000000")
    wname, clr, lv = CASES[case]
    w = bench.WORKLOADS[wname]
    clearance = w["clearance"] if clr is None else clr
    lin_vel = w["lin_vel"] if lv is None else lv
    window, stride = w["window"], w["stride"]
    nsteps = window + 1
    F = (nsteps - 1) // stride + 1
    rm = load_robot(w["robot"])
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    host = bench.workload_batch(env, dict(w, clearance=clearance, lin_vel=lin_vel), N_DRAW, nsteps, seed=0)
    # (the per-substep references are the device expansion of the shipped per-frame ones)
    from ppr_diffphys_b200 import RefsFromFrames
    dev_refs = RefsFromFrames.apply(host["ref_frames"].cuda(), stride, nsteps)
    assert (dev_refs.cpu() - host["refs"]).abs().max() <= 1e-6     # (FMA contraction on the device, none on the host)
    pick = torch.linspace(0, N_DRAW - 1, N_PICK).round().long()
    nb, nq, nqd = rm.nb, rm.nq, rm.nqd
    q_init = host["q_init"].view(N_DRAW, nq)[pick].contiguous()
    qd_init = host["qd_init"].view(N_DRAW, nqd)[pick].contiguous()
    refs = host["refs"].view(nsteps, N_DRAW, nqd)[:, pick].contiguous()
    bs = N_PICK
    caller = Caller(env, bs, nsteps, stride)
    t32 = lambda x: torch.as_tensor(x, dtype=torch.float32, device=dev)
    nI = t32(rm.norm_body_inertia)

    def cuda_call(shared):
        """bench.py step(): shared=True is its default convention (un-replicated parameters), False the reference's
        literal one (--replicate-params, dp_model.py:723-730); torques / res_f = None in both."""
        leaf = lambda t: t.clone().requires_grad_(True)
        rep = (lambda t: t) if shared else (lambda t: t[None].expand(bs, *t.shape).reshape(bs * t.shape[0], *t.shape[1:]).contiguous())
        m = rep(t32(rm.body_mass))
        I0 = rep(nI) * m[:, None, None]
        a = dict(q_init=leaf(q_init.reshape(-1).to(dev)), qd_init=leaf(qd_init.reshape(-1).to(dev)),
                 refs=leaf(refs.reshape(nsteps, -1).to(dev)), target_ke=leaf(rep(t32(rm.joint_target_ke))),
                 target_kd=leaf(rep(t32(rm.joint_target_kd))), body_inv_mass=leaf(1.0 / m), body_inertia=leaf(I0),
                 body_inv_inertia=leaf(torch.linalg.inv(I0)))
        pos, vel = ForwardWarp.apply(a["q_init"], a["qd_init"], None, None, a["refs"], a["target_ke"], a["target_kd"], m,
                                     a["body_inv_mass"], a["body_inertia"], a["body_inv_inertia"], caller)
        bench_loss(pos, vel).backward()
        return pos.detach(), {k: v.grad for k, v in a.items()}

    pos_r, g_rep = cuda_call(shared=False)
    pos_s, g_sh = cuda_call(shared=True)
    # ---- oracle, float64 (and float32 for the noise floor), per-env replicated parameters
    t64 = lambda x: torch.as_tensor(x, dtype=torch.float64)
    o_mass, o_nI = t64(rm.body_mass), t64(rm.norm_body_inertia)
    rep64 = lambda t: t[None].expand(bs, *t.shape).contiguous()
    d = dict(q_init=q_init.double(), qd_init=qd_init.double(), torques=torch.zeros(nsteps, bs, nqd, dtype=torch.float64),
             res_f=torch.zeros(nsteps, bs, nb, 6, dtype=torch.float64), refs=refs.double(),
             target_ke=rep64(t64(rm.joint_target_ke)), target_kd=rep64(t64(rm.joint_target_kd)),
             body_inv_mass=rep64(1.0 / o_mass), body_inertia=rep64(o_nI * o_mass[:, None, None]),
             body_inv_inertia=rep64(torch.linalg.inv(o_nI * o_mass[:, None, None])))
    keys = ["q_init", "qd_init", "refs", "target_ke", "target_kd", "body_inv_mass", "body_inertia", "body_inv_inertia"]
    floor, g64 = fp32_noise_floor(rm, d, stride, F, loss_fn=bench_loss, keys=keys)
    opos, _, _ = oracle_rollout_grads(rm, d, stride, F, loss_fn=bench_loss, keys=keys)
    perr = float((pos_r.cpu().double().reshape(F, bs, nb, 7) - opos).abs().max())
    # ---- per environment: the contact model is only piecewise smooth (contact on/off, stick/slide, +-500 N and +-10 m/s
    # clamps), and an environment that sits ON a switch gets the one-sided derivative of whichever side its rounding
    # lands on (the approximate MUFU div / sqrt of this build vs IEEE: either is a valid fp32 evaluation).  So parity is
    # asserted per environment, at most one of the 16 may be on a switch, and it is still bounded.
    env_of = dict(q_init=lambda t: t.reshape(bs, -1), qd_init=lambda t: t.reshape(bs, -1),
                  refs=lambda t: t.reshape(nsteps, bs, -1).transpose(0, 1).reshape(bs, -1))
    per_env = {k: [rel_err((env_of.get(k, lambda t: t.reshape(bs, -1)))(g_rep[k])[e],
                           (env_of.get(k, lambda t: t.reshape(bs, -1)))(g64[k])[e]) for e in range(bs)] for k in keys}
    tol = {k: max(GRAD_RTOL, 2.0 * floor[k]) for k in keys}
    on_switch = sorted({e for k in keys for e in range(bs) if not per_env[k][e] <= tol[k]})
    worst_regular = {k: max(per_env[k][e] for e in range(bs) if e not in on_switch) for k in keys}
    worst_switch = max([per_env[k][e] for k in keys for e in on_switch], default=0.0)
    print("\n[%s | %s] pose err %.1e; worst regular env: " % (case, layout, perr)
          + "; ".join("%s %.1e (tol %.1e)" % (k, worst_regular[k], tol[k]) for k in keys)
          + "; envs on a branch switch: %s (worst %.1e)" % (on_switch, worst_switch))
    assert perr <= POS_TOL, perr
    assert all(bool(torch.isfinite(g).all()) for g in g_rep.values())
    assert len(on_switch) <= 1 and worst_switch <= 5e-2, (on_switch, worst_switch)
    # ---- bench.py's shared-parameter convention == the replicated one: same trajectories, same per-env gradients, and
    # the gradients of the shared parameters are the sums over environments (dp_model.py:723-725's repeat() backward)
    assert torch.equal(pos_s, pos_r)
    for k in ("q_init", "qd_init", "refs"):
        assert rel_err(g_sh[k], g_rep[k]) <= 1e-6, k
    for k in ("target_ke", "target_kd", "body_inv_mass", "body_inertia", "body_inv_inertia"):
        summed = g_rep[k].reshape(bs, *g_sh[k].shape).sum(0)
        assert rel_err(g_sh[k], summed) <= 1e-5, (k, rel_err(g_sh[k], summed))


@pytest.mark.parametrize("robot", ["human", "laikago"])
def test_cuda_gradients_against_finite_differences(robot):
    """north_star: gradients 'cross-checked by finite differences' -- central differences THROUGH THE CUDA FORWARD KERNEL
    itself (no oracle involved) along random directions of every differentiable input, against <adjoint, direction>.
    The loss is accumulated in float64 from the fp32 outputs; steps are sized so that the loss moves by ~1e-3 of itself
    (far above the fp32 rounding of the rollout, far below its curvature)."""
    from helpers import make_inputs, settle_height
    from ppr_diffphys_b200 import ForwardWarp, SimEnv
    stride, F, bs = 16, 3, 4
    T = stride * (F - 1) + 1
    rm, d = make_inputs(robot, bs=bs, T=T, seed=3, lin_vel=0.3, ang=0.3)
    d = settle_height(rm, d, 0.002)
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    g = torch.Generator().manual_seed(11)
    wp = torch.randn(F, bs * rm.nb, 7, generator=g, dtype=torch.float64)
    wv = torch.randn(F, bs * rm.nb, 6, generator=g, dtype=torch.float64) * 0.1
    flat = dict(q_init=d["q_init"].reshape(-1), qd_init=d["qd_init"].reshape(-1), refs=d["refs"].reshape(T, -1),
                target_ke=d["target_ke"].reshape(-1), target_kd=d["target_kd"].reshape(-1),
                body_inv_mass=d["body_inv_mass"].reshape(-1), body_inertia=d["body_inertia"].reshape(-1, 3, 3),
                body_inv_inertia=d["body_inv_inertia"].reshape(-1, 3, 3))

    def run(x, need_grad):
        a = {k: v.to(dev, torch.float32).requires_grad_(need_grad) for k, v in x.items()}
        pos, vel = ForwardWarp.apply(a["q_init"], a["qd_init"], None, None, a["refs"], a["target_ke"], a["target_kd"],
                                     (1.0 / a["body_inv_mass"]).detach(), a["body_inv_mass"], a["body_inertia"],
                                     a["body_inv_inertia"], Caller(env, bs, T, stride))
        if need_grad:
            torch.autograd.backward([pos, vel], [wp.to(dev, torch.float32), wv.to(dev, torch.float32)])
        loss = float((pos.detach().cpu().double() * wp).sum() + (vel.detach().cpu().double() * wv).sum())
        return loss, a

    L0, a0 = run(flat, True)
    # direction per input; the step is sized from the analytic directional derivative so that the loss moves by ~2e-3
    # of its magnitude (bounded per input to stay in the linear regime)
    max_h = dict(q_init=1e-4, qd_init=1e-2, refs=2e-3, target_ke=5.0, target_kd=0.1, body_inv_mass=3e-2,
                 body_inertia=3e-2, body_inv_inertia=3e-2)
    worst, detail = {}, []
    for k in max_h:
        x0 = flat[k]
        u = torch.randn(x0.shape, generator=g, dtype=torch.float64)
        if k in ("body_inv_mass", "body_inertia", "body_inv_inertia"):
            u = u * x0.abs()               # relative perturbation
        if k == "q_init":
            # joint angles only: moving the ROOT pose walks every foot through contact switches, and the step needed to
            # stay between two of them (<= 1e-5, measured with the float64 port) is below fp32 resolution; the root
            # components are covered by the oracle comparisons above and by the float64 FD checks of tests/test_oracle.py
            u = u.view(bs, -1); u[:, :7] = 0; u = u.reshape(-1)
        if k in ("target_ke", "target_kd"):
            u = u.view(bs, -1); u[:, :6] = 0; u = u.reshape(-1)       # no PD on the six root dofs
        an = float((a0[k].grad.detach().cpu().double() * u).sum())
        h = min(max_h[k], 2e-3 * max(abs(L0), 1.0) / max(abs(an), 1e-12))
        lp, _ = run(dict(flat, **{k: x0 + h * u}), False)
        lm, _ = run(dict(flat, **{k: x0 - h * u}), False)
        fd = (lp - lm) / (2 * h)
        worst[k] = abs(fd - an) / max(abs(fd), abs(an), 1e-12)
        detail.append("%s: fd %.5e an %.5e h %.1e" % (k, fd, an, h))
    print("\n[finite differences through the CUDA kernels, %s] loss %.4f; " % (robot, L0)
          + "; ".join("%s %.1e" % kv for kv in worst.items()) + " | " + "; ".join(detail))
    for k, e in worst.items():
        assert e <= 3e-2, (k, e)
