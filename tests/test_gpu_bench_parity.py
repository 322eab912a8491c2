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
from helpers import fp32_noise_floor, rel_err  # noqa: E402

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
    from ppr_diffphys_b200.synth import make_batch
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
    host = make_batch(env, N_DRAW, nsteps, seed=0, clearance=clearance, lin_vel=lin_vel)   # bench.py:194
    pick = torch.linspace(0, N_DRAW - 1, N_PICK).round().long()
    nb, nq, nqd = rm.nb, rm.nq, rm.nqd
    q_init = host["q_init"].view(N_DRAW, nq)[pick].contiguous()
    qd_init = host["qd_init"].view(N_DRAW, nqd)[pick].contiguous()
    refs = host["refs"].view(nsteps, N_DRAW, nqd)[:, pick].contiguous()
    # ---- CUDA, the call of bench.py step() (shared parameters, null torques / res_f)
    nI = torch.as_tensor(rm.norm_body_inertia, device=dev)
    p_ke = torch.as_tensor(rm.joint_target_ke, device=dev).clone().requires_grad_(True)
    p_kd = torch.as_tensor(rm.joint_target_kd, device=dev).clone().requires_grad_(True)
    p_mass = torch.as_tensor(rm.body_mass, device=dev).clone().requires_grad_(True)
    c_q = q_init.reshape(-1).to(dev).requires_grad_(True)
    c_qd = qd_init.reshape(-1).to(dev).requires_grad_(True)
    c_refs = refs.reshape(nsteps, -1).to(dev).requires_grad_(True)
    inv_m, I = 1.0 / p_mass, nI * p_mass[:, None, None]
    inv_I = torch.linalg.inv(nI) * inv_m[:, None, None]
    pos, vel = ForwardWarp.apply(c_q, c_qd, None, None, c_refs, p_ke, p_kd, p_mass, inv_m, I, inv_I,
                                 Caller(env, N_PICK, nsteps, stride))
    bench_loss(pos, vel).backward()
    # ---- oracle, float64 (and float32 for the noise floor), per-env replicated parameters chained to the shared ones
    bs = N_PICK
    o_ke = torch.as_tensor(rm.joint_target_ke, dtype=torch.float64)
    o_kd = torch.as_tensor(rm.joint_target_kd, dtype=torch.float64)
    o_mass = torch.as_tensor(rm.body_mass, dtype=torch.float64)
    o_nI = torch.as_tensor(rm.norm_body_inertia, dtype=torch.float64)
    rep = lambda t: t[None].expand(bs, *t.shape).contiguous()
    d = dict(q_init=q_init.double(), qd_init=qd_init.double(), torques=torch.zeros(nsteps, bs, nqd, dtype=torch.float64),
             res_f=torch.zeros(nsteps, bs, nb, 6, dtype=torch.float64), refs=refs.double(), target_ke=rep(o_ke),
             target_kd=rep(o_kd), body_inv_mass=rep(1.0 / o_mass), body_inertia=rep(o_nI * o_mass[:, None, None]),
             body_inv_inertia=rep(torch.linalg.inv(o_nI * o_mass[:, None, None])))
    from helpers import oracle_rollout_grads
    keys = ["q_init", "qd_init", "refs", "target_ke", "target_kd", "body_inv_mass", "body_inertia", "body_inv_inertia"]
    floor, g64 = fp32_noise_floor(rm, d, stride, F, loss_fn=bench_loss, keys=keys)
    opos, _, _ = oracle_rollout_grads(rm, d, stride, F, loss_fn=bench_loss, keys=keys)
    perr = float((pos.detach().cpu().double().reshape(F, bs, nb, 7) - opos).abs().max())
    # the feet really touch the ground in this window (the workload is about contacts)
    assert float(opos[-1][..., 1].min()) < 0.2
    # chain of dp_model.py:725-730 for the mass-related gradients, and the sum over envs of the shared ones
    m64 = o_mass.clone().requires_grad_(True)
    inv_m64 = 1.0 / m64
    I64 = o_nI * m64[:, None, None]
    invI64 = torch.linalg.inv(I64)
    (g_mass,) = torch.autograd.grad((inv_m64 * g64["body_inv_mass"].sum(0)).sum() + (I64 * g64["body_inertia"].sum(0)).sum()
                                    + (invI64 * g64["body_inv_inertia"].sum(0)).sum(), m64)
    mass_floor = max(floor["body_inv_mass"], floor["body_inertia"], floor["body_inv_inertia"])
    got = {"q_init": (c_q.grad, g64["q_init"], floor["q_init"]), "qd_init": (c_qd.grad, g64["qd_init"], floor["qd_init"]),
           "refs": (c_refs.grad, g64["refs"], floor["refs"]),
           "target_ke": (p_ke.grad, g64["target_ke"].sum(0), floor["target_ke"]),
           "target_kd": (p_kd.grad, g64["target_kd"].sum(0), floor["target_kd"]),
           "body_mass": (p_mass.grad, g_mass, mass_floor)}
    report, bad = [], []
    for k, (g, ref, fl) in got.items():
        e = rel_err(g, ref)
        tol = max(GRAD_RTOL, 2.0 * fl)
        report.append("%s err %.1e (fp32 floor %.1e, tol %.1e)" % (k, e, fl, tol))
        if not (e <= tol) or not bool(torch.isfinite(g).all()):
            bad.append(k)
    print("\n[%s | %s] pose err %.1e; " % (case, layout, perr) + "; ".join(report))
    assert perr <= POS_TOL, perr
    assert not bad, (bad, report)


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
    # perturbation scale per input (units of that input)
    scale = dict(q_init=2e-4, qd_init=2e-3, refs=5e-4, target_ke=0.5, target_kd=0.02, body_inv_mass=1e-3,
                 body_inertia=None, body_inv_inertia=None)
    worst = {}
    for k, h in scale.items():
        x0 = flat[k]
        u = torch.randn(x0.shape, generator=g, dtype=torch.float64)
        if h is None:                      # relative perturbation of the (symmetric positive) inertia tensors
            u, h = u * x0.abs(), 1e-3
        if k == "q_init":                  # keep the root quaternion direction generic but small
            u = u.view(bs, -1); u[:, :3] *= 0.1; u = u.reshape(-1)
        if k in ("target_ke", "target_kd"):
            u = u.view(bs, -1); u[:, :6] = 0; u = u.reshape(-1)       # no PD on the six root dofs
        lp, _ = run(dict(flat, **{k: x0 + h * u}), False)
        lm, _ = run(dict(flat, **{k: x0 - h * u}), False)
        fd = (lp - lm) / (2 * h)
        an = float((a0[k].grad.detach().cpu().double() * u).sum())
        worst[k] = abs(fd - an) / max(abs(fd), abs(an), 1e-12)
    print("\n[finite differences through the CUDA kernels, %s] loss %.4f; " % (robot, L0)
          + "; ".join("%s %.1e" % kv for kv in worst.items()))
    for k, e in worst.items():
        assert e <= 3e-2, (k, e)
