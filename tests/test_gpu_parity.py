"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes) behind the reference-shaped
autograd Functions, against (a) the committed golden vectors of the float64 oracle, (b) the oracle run live on
seeded inputs, (c) size-independent properties at BASELINE.json's full sizes, (d) edge cases.

Tolerances (north_star): body pose after a 64-substep window <= 1e-4 (m / quaternion component ~ rad);
every gradient within relative error 1e-3 (norm-wise).  One documented exception, laikago in stiff contact, where
1e-3 is below what single precision resolves: there the bound is max(1e-3, 2 x the fp32 noise floor of that
gradient), the floor being measured, not assumed (helpers.fp32_noise_floor: the float64 oracle re-run in float32 on
the same inputs; tests/test_oracle.py pins its range on the CPU).  Achieved errors are printed with pytest -s."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from helpers import fp32_noise_floor, make_inputs, make_mixed_robot, settle_height

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEYS = ["q_init", "qd_init", "torques", "res_f", "refs", "target_ke", "target_kd", "body_inv_mass", "body_inertia",
        "body_inv_inertia"]
POS_TOL, GRAD_RTOL = 1e-4, 1e-3
# laikago IN CONTACT is ill-conditioned in fp32: 0.16 kg lower legs on up to 96 penalty contacts of 1e4 N/m each sit
# near the stability limit of semi-implicit Euler at dt = 5e-4 (k dt^2 / m ~ 1), so rounding differences are
# amplified ~1e3x over a 64-substep window: an independent float32 evaluation of the reference formulation (the
# autograd oracle in float32) differs from float64 by 1e-3 .. 5e-3 per gradient on the committed fixture.  The
# tolerance for that robot is therefore per gradient max(GRAD_RTOL, 2 x measured floor); the airborne laikago fixture
# and human / quad in contact hold the strict 1e-3.


def grad_tolerances(rm, d, stride, F, **kw):
    """{key: max(1e-3, 2 x fp32 noise floor)} for the inputs ``d`` ([bs,...] layout)."""
    floor, _ = fp32_noise_floor(rm, d, stride, F, **kw)
    return {k: max(GRAD_RTOL, 2.0 * v) for k, v in floor.items()}, floor


@pytest.fixture(autouse=True, params=["throughput-layout", "latency-layout"])
def layout(request, monkeypatch):
    """Every test of this file runs under both environment packings: the block / warp packing used for batches
    that fill the GPU, and the one-environment-per-warp layout small batches get by default (PPR_LATENCY_ENVS is
    read by ppr_model_create)."""
    monkeypatch.setenv("PPR_LATENCY_ENVS", "0" if request.param == "throughput-layout" else "1000000")
    return request.param


class Caller:
    """Stand-in for the reference's ``phys_model`` object handed to ForwardWarp.apply (dp_model.py:733-746)."""

    def __init__(self, env, num_envs, nsteps, stride, dt=5e-4):
        self.env, self.num_envs, self.dt = env, num_envs, dt
        self.steps_idx = range(nsteps)
        self.frame2step = [i for i in range(nsteps) if i % stride == 0]


def flat_args(d, dev, requires_grad=True, drop=()):
    """[bs,...] oracle layout -> the reference's flattened call layout (dp_model.py:563-572,697-699,723-730)."""
    bs = d["q_init"].shape[0]
    T = d["refs"].shape[0]
    t = lambda x: x.to(dev, torch.float32)
    a = dict(q_init=t(d["q_init"]).reshape(-1), qd_init=t(d["qd_init"]).reshape(-1),
             torques=t(d["torques"]).reshape(T, -1), res_f=t(d["res_f"]).reshape(T, -1, 6),
             refs=t(d["refs"]).reshape(T, -1), target_ke=t(d["target_ke"]).reshape(-1),
             target_kd=t(d["target_kd"]).reshape(-1), body_inv_mass=t(d["body_inv_mass"]).reshape(-1),
             body_inertia=t(d["body_inertia"]).reshape(-1, 3, 3),
             body_inv_inertia=t(d["body_inv_inertia"]).reshape(-1, 3, 3))
    for k in drop:
        a[k] = None
    if requires_grad:
        for k, v in a.items():
            if v is not None:
                v.requires_grad_(True)
    a["body_mass"] = (1.0 / a["body_inv_mass"]).detach()
    return a, bs, T


def run_cuda(env, a, bs, T, stride):
    from ppr_diffphys_b200 import ForwardWarp
    caller = Caller(env, bs, T, stride)
    pos, vel = ForwardWarp.apply(a["q_init"], a["qd_init"], a["torques"], a["res_f"], a["refs"], a["target_ke"],
                                 a["target_kd"], a["body_mass"], a["body_inv_mass"], a["body_inertia"],
                                 a["body_inv_inertia"], caller)
    return pos, vel, caller


def rel(a, b):
    return float((a.double().cpu().reshape(-1) - b.double().reshape(-1)).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("fixture", ["laikago", "laikago_air", "human", "quad"])
def test_golden_forward_and_gradients(fixture):
    from ppr_diffphys_b200 import SimEnv
    z = np.load(os.path.join(GOLDEN, "rollout_%s.npz" % fixture))
    robot = str(z["robot"])
    dev = torch.device("cuda:0")
    env = SimEnv(robot)
    d = {k: torch.from_numpy(z["in_" + k]) for k in KEYS}
    gtol, floor = {k: GRAD_RTOL for k in KEYS}, {k: 0.0 for k in KEYS}
    if fixture == "laikago":
        gtol, floor = grad_tolerances(env.model, d, int(z["stride"]), int(z["nframes"]),
                                      adj_pos=torch.from_numpy(z["adj_pos"]), adj_vel=torch.from_numpy(z["adj_vel"]))
    a, bs, T = flat_args(d, dev)
    stride, F = int(z["stride"]), int(z["nframes"])
    pos, vel, caller = run_cuda(env, a, bs, T, stride)
    gpos = torch.from_numpy(z["pos"]).reshape(F, -1, 7)
    gvel = torch.from_numpy(z["vel"]).reshape(F, -1, 6)
    assert (pos.cpu().double() - gpos).abs().max() <= POS_TOL
    assert (vel.cpu().double() - gvel).abs().max() <= 5e-3  # velocities ~O(1-10) m/s, rad/s
    assert rel(torch.stack(caller.grfs), torch.from_numpy(z["grf"])) <= 1e-3
    assert rel(torch.stack(caller.jafs), torch.from_numpy(z["jaf"])) <= 1e-3
    assert np.allclose(np.stack(list(caller.sim_trajs)), z["pos"][:, 0], atol=POS_TOL)
    adj_pos = torch.from_numpy(z["adj_pos"]).reshape(F, -1, 7).to(dev, torch.float32)
    adj_vel = torch.from_numpy(z["adj_vel"]).reshape(F, -1, 6).to(dev, torch.float32)
    torch.autograd.backward([pos, vel], [adj_pos, adj_vel])
    errs = {}
    for k in KEYS:
        g = a[k].grad
        assert g is not None and torch.isfinite(g).all(), k
        errs[k] = rel(g, torch.from_numpy(z["grad_" + k]))
    print("\n[golden %s] " % fixture + "; ".join("%s %.1e (floor %.1e)" % (k, errs[k], floor[k]) for k in KEYS))
    for k in KEYS:
        assert errs[k] <= gtol[k], (k, errs[k], gtol[k])


@pytest.mark.parametrize("robot,bs", [("laikago", 5), ("human", 2), ("quad", 3)])
def test_live_oracle_parity_null_forces(robot, bs):
    """torques / res_f passed as None (fast path) == oracle with exact zeros; odd batch sizes exercise the
    partially filled last warp."""
    from oracle import sim_oracle as so
    from ppr_diffphys_b200 import SimEnv
    stride, F = 16, 3
    T = stride * (F - 1) + 1
    rm, d = make_inputs(robot, bs=bs, T=T, seed=31, lin_vel=0.3, ang=0.3)
    d = settle_height(rm, d, 0.002)
    d = {k: v.float().double() for k, v in d.items()}
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    a, _, _ = flat_args(d, dev, drop=("torques", "res_f"))
    pos, vel, _ = run_cuda(env, a, bs, T, stride)
    m = so.OracleModel(rm)
    o = {k: d[k].clone().requires_grad_(True) for k in KEYS}
    opos, ovel = so.rollout(m, o["q_init"], o["qd_init"], d["torques"] * 0, d["res_f"] * 0, o["refs"], o["target_ke"],
                            o["target_kd"], o["body_inv_mass"], o["body_inertia"], o["body_inv_inertia"], 5e-4, stride,
                            F)[:2]
    assert (pos.cpu().double() - opos.detach().reshape(F, -1, 7)).abs().max() <= POS_TOL
    loss_c = (pos ** 2).sum() + 0.1 * (vel ** 2).sum()
    loss_o = (opos ** 2).sum() + 0.1 * (ovel ** 2).sum()
    loss_c.backward()
    keys = [k for k in KEYS if k not in ("torques", "res_f")]
    grads = torch.autograd.grad(loss_o, [o[k] for k in keys])
    tol = {k: GRAD_RTOL for k in keys}
    if robot == "laikago":
        d0 = dict(d, torques=d["torques"] * 0, res_f=d["res_f"] * 0)
        tol, _ = grad_tolerances(rm, d0, stride, F, loss_fn=lambda p, v: (p ** 2).sum() + 0.1 * (v ** 2).sum(), keys=keys)
    for k, g in zip(keys, grads):
        assert rel(a[k].grad, g) <= tol[k], (k, rel(a[k].grad, g), tol[k])


def test_generic_kernel_instance_mixed_features():
    """The generic instance (JM_ALL, LIMITS, QOFF): FIXED + REVOLUTE + COMPOUND joints, active limit springs,
    non-identity joint_X_c, sphere / capsule contacts with thickness, two materials with kd > 0, torques and res_f."""
    from oracle import sim_oracle as so
    from ppr_diffphys_b200 import SimEnv
    rm = make_mixed_robot()
    stride, F, bs = 16, 3, 5
    T = stride * (F - 1) + 1
    rm, d = make_inputs(rm, bs=bs, T=T, seed=17, ang=0.25, res_f_std=0.05, torque_std=0.05, lin_vel=0.3, qd_std=0.05)
    d = settle_height(rm, d, 0.004)
    d = {k: v.float().double() for k, v in d.items()}
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    a, _, _ = flat_args(d, dev)
    pos, vel, caller = run_cuda(env, a, bs, T, stride)
    m = so.OracleModel(rm)
    o = {k: d[k].clone().requires_grad_(True) for k in KEYS}
    opos, ovel, ogrf, ojaf = so.rollout(m, o["q_init"], o["qd_init"], o["torques"], o["res_f"], o["refs"],
                                        o["target_ke"], o["target_kd"], o["body_inv_mass"], o["body_inertia"],
                                        o["body_inv_inertia"], 5e-4, stride, F)
    assert ogrf.abs().max() > 1.0
    assert (pos.cpu().double() - opos.detach().reshape(F, -1, 7)).abs().max() <= POS_TOL
    assert rel(torch.stack(caller.grfs), ogrf) <= 1e-3 and rel(torch.stack(caller.jafs), ojaf) <= 1e-3
    g = torch.Generator().manual_seed(2)
    wp = torch.randn(opos.shape, generator=g, dtype=torch.float64)
    wv = torch.randn(ovel.shape, generator=g, dtype=torch.float64) * 0.1
    torch.autograd.backward([pos, vel], [wp.reshape(F, -1, 7).to(dev, torch.float32),
                                         wv.reshape(F, -1, 6).to(dev, torch.float32)])
    grads = torch.autograd.grad((opos * wp).sum() + (ovel * wv).sum(), [o[k] for k in KEYS])
    for k, gr in zip(KEYS, grads):
        assert rel(a[k].grad, gr) <= GRAD_RTOL, (k, rel(a[k].grad, gr))
    # FK of the generic instance
    bq, bqd = env.fk(a["q_init"].detach().view(bs, -1), a["qd_init"].detach().view(bs, -1))
    obq, obqd = so.eval_fk(m, d["q_init"], d["qd_init"])
    assert (bq.cpu().double() - obq).abs().max() < 2e-6 and (bqd.cpu().double() - obqd).abs().max() < 2e-5


@pytest.mark.parametrize("robot", ["laikago", "human", "quad"])
def test_fk_parity_and_grad(robot):
    from oracle import sim_oracle as so
    from ppr_diffphys_b200 import ForwardKinematics, SimEnv
    T, bs = 4, 3
    rm, d = make_inputs(robot, bs=T * bs, T=1, seed=5, ang=0.8, qd_std=0.5, quat_noise=0.2, normalize_quat=False)
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    q = d["q_init"].float().view(T, bs, -1).to(dev).requires_grad_(True)
    qd = d["qd_init"].float().view(T, bs, -1).to(dev).requires_grad_(True)
    bq, bqd, frames = ForwardKinematics.apply(q, qd, env)
    assert bq.shape == (bs, T, rm.nb, 7) and bqd.shape == (bs, T, rm.nb, 6) and len(frames) == T
    m = so.OracleModel(rm)
    oq = d["q_init"].float().double().requires_grad_(True)
    oqd = d["qd_init"].float().double().requires_grad_(True)
    obq, obqd = so.eval_fk(m, oq, oqd)
    obq_r = obq.view(T, bs, rm.nb, 7).permute(1, 0, 2, 3)
    obqd_r = obqd.view(T, bs, rm.nb, 6).permute(1, 0, 2, 3)
    assert (bq.cpu().double() - obq_r.detach()).abs().max() < 2e-6
    assert (bqd.cpu().double() - obqd_r.detach()).abs().max() < 2e-5
    g = torch.Generator().manual_seed(3)
    w1 = torch.randn(bq.shape, generator=g) * 0.01
    w2 = torch.randn(bqd.shape, generator=g) * 0.01
    ((bq * w1.to(dev)).sum() + (bqd * w2.to(dev)).sum()).backward()
    gq, gqd = torch.autograd.grad((obq_r * w1.double()).sum() + (obqd_r * w2.double()).sum(), [oq, oqd])
    # reference post-processing: upper clamp at +1 (dp_model.py:1110,1123) -- weights chosen so it is inactive
    assert gq.max() < 1 and gqd.max() < 1
    assert rel(q.grad, gq.view(T, bs, -1)) < 1e-4
    assert rel(qd.grad, gqd.view(T, bs, -1)) < 1e-4


# ------------------------------------------------------------------ properties at full size (BASELINE.json configs)
@pytest.mark.parametrize("robot,bs", [("quad", 1024), ("human", 4096), ("laikago", 4096)])
def test_full_size_properties(robot, bs):
    """Config 3/4/5 sizes, 64-substep window: (1) bit-exact determinism (no atomics), (2) batch independence: the
    first 7 envs computed alone give bit-identical trajectories and gradients, (3) backward is linear in the incoming
    adjoints, (4) all outputs finite."""
    from ppr_diffphys_b200 import SimEnv
    stride, F = 32, 3
    T = stride * (F - 1) + 1
    rm, d = make_inputs(robot, bs=bs, T=T, seed=0, lin_vel=1.0 if robot == "human" else 0.0)
    d = settle_height(rm, d, 0.002)
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    env.set_team_envs(0)   # the 7-env sub-batch must take the same layout as the full batch: bit-identity holds per kernel
                           # instance (the team layout has its own test below)

    def run(dd, adj_scale=1.0, seed=1):
        a, n, _ = flat_args(dd, dev, drop=("torques", "res_f"))
        pos, vel, _ = run_cuda(env, a, n, T, stride)
        g = torch.Generator(device="cpu").manual_seed(seed)
        ap = (torch.randn(pos.shape, generator=g) * adj_scale).to(dev)
        av = (torch.randn(vel.shape, generator=g) * adj_scale * 0.1).to(dev)
        torch.autograd.backward([pos, vel], [ap, av])
        return pos.detach(), vel.detach(), {k: a[k].grad for k in KEYS if a.get(k) is not None}, (ap, av), a

    p1, v1, g1, adj, a1 = run(d)
    p2, v2, g2, _, _ = run(d)
    assert torch.equal(p1, p2) and torch.equal(v1, v2)
    for k in g1:
        assert torch.isfinite(g1[k]).all(), k
        assert torch.equal(g1[k], g2[k]), k
    # batch independence
    sub = {k: (v[:, :7] if k in ("torques", "res_f", "refs") else v[:7]) for k, v in d.items()}
    a, n, _ = flat_args(sub, dev, drop=("torques", "res_f"))
    ps, vs, _ = run_cuda(env, a, 7, T, stride)
    nb = rm.nb
    assert torch.equal(ps, p1.view(F, bs, nb, 7)[:, :7].reshape(F, -1, 7))
    ap, av = adj
    torch.autograd.backward([ps, vs], [ap.view(F, bs, nb, 7)[:, :7].reshape(F, -1, 7).contiguous(),
                                       av.view(F, bs, nb, 6)[:, :7].reshape(F, -1, 6).contiguous()])
    assert torch.equal(a["q_init"].grad, g1["q_init"].view(bs, -1)[:7].reshape(-1))
    assert torch.equal(a["refs"].grad, g1["refs"].view(T, bs, -1)[:, :7].reshape(T, -1))
    # linearity of the adjoint in the seeds: bwd(2a) == 2 bwd(a) exactly (power of two)
    _, _, g3, _, _ = run(d, adj_scale=2.0)
    for k in g1:
        assert torch.allclose(g3[k], 2 * g1[k], rtol=1e-5, atol=1e-6 * float(g1[k].abs().max())), k


@pytest.mark.parametrize("robot", ["laikago", "human"])
def test_team_layout_determinism_and_batch_independence(robot):
    """Team layout (one env per block, contacts on helper warps): bit-exact determinism, and the first 7 of 40 envs computed
    alone give bit-identical trajectories and gradients (an environment never sees its neighbours)."""
    from ppr_diffphys_b200 import SimEnv
    stride, F, bs = 16, 3, 40
    T = stride * (F - 1) + 1
    rm, d = make_inputs(robot, bs=bs, T=T, seed=2, lin_vel=0.3)
    d = settle_height(rm, d, 0.002)
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    env.set_latency_envs(1 << 20)
    env.set_team_envs(1 << 20)
    g = torch.Generator(device="cpu").manual_seed(5)
    nb = rm.nb
    ap = torch.randn(F, bs, nb, 7, generator=g).to(dev)
    av = (torch.randn(F, bs, nb, 6, generator=g) * 0.1).to(dev)

    def run(dd, n):
        a, _, _ = flat_args(dd, dev, drop=("torques", "res_f"))
        pos, vel, caller = run_cuda(env, a, n, T, stride)
        torch.autograd.backward([pos, vel], [ap[:, :n].reshape(F, -1, 7).contiguous(), av[:, :n].reshape(F, -1, 6).contiguous()])
        return pos.detach(), vel.detach(), torch.stack(caller.grfs), {k: a[k].grad for k in KEYS if a.get(k) is not None}

    p1, v1, f1, g1 = run(d, bs)
    p2, v2, f2, g2 = run(d, bs)
    assert torch.equal(p1, p2) and torch.equal(v1, v2) and torch.equal(f1, f2)
    assert float(f1.abs().max()) > 0                       # the fixture is in contact: the helpers did something
    for k in g1:
        assert torch.isfinite(g1[k]).all() and torch.equal(g1[k], g2[k]), k
    sub = {k: (v[:, :7] if k in ("torques", "res_f", "refs") else v[:7]) for k, v in d.items()}
    ps, vs, fs, gs = run(sub, 7)
    assert torch.equal(ps, p1.view(F, bs, nb, 7)[:, :7].reshape(F, -1, 7))
    assert torch.equal(fs, f1.view(F, bs, nb, 6)[:, :7].reshape(F, -1, 6))
    assert torch.equal(gs["q_init"], g1["q_init"].view(bs, -1)[:7].reshape(-1))
    assert torch.equal(gs["refs"], g1["refs"].view(T, bs, -1)[:, :7].reshape(T, -1))


def test_free_fall_full_size():
    from ppr_diffphys_b200 import SimEnv
    bs, stride, F = 2048, 32, 3
    T = stride * (F - 1) + 1
    rm, d = make_inputs("laikago", bs=bs, T=T, seed=0, ang=0.3, qd_std=0.0, height=3.0, ref_amp=0.0, quat_noise=0.0)
    d["refs"][:, :, 6:] = d["q_init"][None, :, 7:]
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    a, n, _ = flat_args(d, dev, requires_grad=False, drop=("torques", "res_f"))
    pos, vel, _ = run_cuda(env, a, n, T, stride)
    g = float(rm.gravity[1])
    vy = vel.view(F, bs, rm.nb, 6)[..., 4]
    assert torch.allclose(vy[2], torch.full_like(vy[2], g * 64 * 5e-4), atol=2e-5)
    dy = pos.view(F, bs, rm.nb, 7)[2, ..., 1] - pos.view(F, bs, rm.nb, 7)[0, ..., 1]
    assert torch.allclose(dy, torch.full_like(dy, g * 5e-4 * 5e-4 * 64 * 65 / 2), atol=2e-5)


# ------------------------------------------------------------------ edge cases
def test_zero_forces_equal_null_pointers():
    from ppr_diffphys_b200 import SimEnv
    stride, F, bs = 8, 2, 4
    T = stride * (F - 1) + 1
    rm, d = make_inputs("human", bs=bs, T=T, seed=3)
    d = settle_height(rm, d, 0.003)
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    a0, _, _ = flat_args(d, dev)
    a1, _, _ = flat_args(d, dev, drop=("torques", "res_f"))
    p0, v0, _ = run_cuda(env, a0, bs, T, stride)
    p1, v1, _ = run_cuda(env, a1, bs, T, stride)
    assert torch.equal(p0, p1) and torch.equal(v0, v1)
    (p0.sum() + v0.sum()).backward()
    (p1.sum() + v1.sum()).backward()
    assert torch.equal(a0["refs"].grad, a1["refs"].grad)
    assert a0["torques"].grad.shape == (T, bs * rm.nqd) and a0["res_f"].grad.shape == (T, bs * rm.nb, 6)
    # last substep is never differentiated (dp_model.py:397)
    assert float(a0["refs"].grad[-1].abs().max()) == 0.0 and float(a0["res_f"].grad[-1].abs().max()) == 0.0
    # the 6 root dofs carry no PD torque
    assert float(a0["refs"].grad.view(T, bs, -1)[..., :6].abs().max()) == 0.0


@pytest.mark.parametrize("robot", ["laikago", "human"])
def test_shared_parameters_equal_replicated(robot):
    """Un-replicated [nqd]/[nb]/[nb,3,3] parameters == the reference's per-env replication (dp_model.py:723-730):
    bit-identical trajectories, and gradients equal to the sum over envs of the per-env gradients."""
    from ppr_diffphys_b200 import ForwardWarp, SimEnv, load_robot
    stride, F, bs = 16, 3, 6
    T = stride * (F - 1) + 1
    rm, d = make_inputs(robot, bs=bs, T=T, seed=12)
    d = settle_height(rm, d, 0.002)
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    ke = torch.as_tensor(rm.joint_target_ke, device=dev)
    kd = torch.as_tensor(rm.joint_target_kd, device=dev)
    mass = torch.as_tensor(rm.body_mass, device=dev) * 1.1
    nI = torch.as_tensor(rm.norm_body_inertia, device=dev)
    caller = Caller(env, bs, T, stride)
    base = dict(q=d["q_init"].float().reshape(-1).to(dev), qd=d["qd_init"].float().reshape(-1).to(dev),
                refs=d["refs"].float().reshape(T, -1).to(dev))

    def run(shared):
        p = [x.clone().requires_grad_(True) for x in (ke, kd, 1.0 / mass, nI * mass[:, None, None],
                                                      torch.linalg.inv(nI * mass[:, None, None]))]
        if shared:
            args = p
        else:
            args = [x[None].expand(bs, *x.shape).reshape(bs * x.shape[0], *x.shape[1:]) for x in p]
        pos, vel = ForwardWarp.apply(base["q"], base["qd"], None, None, base["refs"], args[0], args[1],
                                     mass if shared else mass.repeat(bs), args[2], args[3], args[4], caller)
        ((pos ** 2).sum() + (vel ** 2).sum() * 0.1).backward()
        return pos.detach(), [x.grad for x in p]

    p0, g0 = run(False)
    p1, g1 = run(True)
    assert torch.equal(p0, p1)
    for a, b in zip(g0, g1):
        assert a.shape == b.shape and torch.allclose(a, b, rtol=1e-4, atol=1e-6 * float(a.abs().max()))


@pytest.mark.parametrize("robot,every", [("laikago", 4), ("laikago", 7), ("human", 16), ("quad", 5), ("mixed", 3)])
def test_checkpoint_every_k_recompute(robot, every):
    """Checkpoint policy (ppr_model_set_checkpoint_every): keeping the state every K substeps and re-computing the
    rest inside the adjoint gives the same trajectories bit for bit and the same gradients as K = 1, with a K-fold
    smaller stored checkpoint.  K = 7 / 5 / 3 do not divide the 33 / 65 substeps (ragged last segment)."""
    from ppr_diffphys_b200 import SimEnv
    stride, F, bs = 16, 3, 9
    T = stride * (F - 1) + 1
    if robot == "mixed":
        rm, d = make_inputs(make_mixed_robot(), bs=bs, T=T, seed=4, ang=0.25, res_f_std=0.05, torque_std=0.05,
                            lin_vel=0.3, qd_std=0.05)
        d = settle_height(rm, d, 0.004)
    else:
        rm, d = make_inputs(robot, bs=bs, T=T, seed=21)
        d = settle_height(rm, d, 0.002)
    dev = torch.device("cuda:0")
    out = {}
    for K in (1, every):
        env = SimEnv(rm)
        env.set_team_envs(0)   # K > 1 exists in the warp / block layouts only: compare within one kernel family (bit for bit)
        ws1 = env._lib.ppr_rollout_workspace_bytes(env._h, bs, T)
        env.set_checkpoint_every(K)
        ws = env._lib.ppr_rollout_workspace_bytes(env._h, bs, T)
        assert ws * T == ws1 * (-(-T // K) + (K if K > 1 else 0))
        a, _, _ = flat_args(d, dev)
        pos, vel, _ = run_cuda(env, a, bs, T, stride)
        w = torch.linspace(0.5, 1.5, pos.numel(), device=dev).reshape(pos.shape)
        ((pos * w).sum() + (vel ** 2).sum() * 0.01).backward()
        out[K] = (pos.detach(), vel.detach(), {k: a[k].grad for k in KEYS})
    assert torch.equal(out[1][0], out[every][0]) and torch.equal(out[1][1], out[every][1])
    for k in KEYS:
        g1, gk = out[1][2][k], out[every][2][k]
        assert torch.isfinite(gk).all()
        assert rel(gk, g1.double().cpu()) < 1e-4, k
    with pytest.raises(Exception):
        env.set_checkpoint_every(0)


@pytest.mark.parametrize("robot", ["laikago", "human", "quad", "mixed"])
def test_latency_layout_equals_throughput_layout(robot):
    """Small batches run one environment per warp (ppr_model_set_latency_envs) or, smaller still, one per block of three
    warps with the ground contacts on two helper warps (ppr_model_set_team_envs); large ones the block / warp packing
    chosen for throughput: same arithmetic per body, so trajectories, force side channels and gradients agree to
    rounding."""
    from ppr_diffphys_b200 import SimEnv
    stride, F, bs = 16, 3, 11
    T = stride * (F - 1) + 1
    if robot == "mixed":
        rm, d = make_inputs(make_mixed_robot(), bs=bs, T=T, seed=4, ang=0.25, res_f_std=0.05, torque_std=0.05,
                            lin_vel=0.3, qd_std=0.05)
        d = settle_height(rm, d, 0.004)
    else:
        rm, d = make_inputs(robot, bs=bs, T=T, seed=33)
        d = settle_height(rm, d, 0.002)
    dev = torch.device("cuda:0")
    out = []
    for lat, team in ((0, 0), (1 << 20, 0), (1 << 20, 1 << 20)):
        env = SimEnv(rm)
        env.set_latency_envs(lat)
        env.set_team_envs(team)
        a, _, _ = flat_args(d, dev)
        pos, vel, caller = run_cuda(env, a, bs, T, stride)
        ((pos ** 2).sum() + (vel ** 2).sum() * 0.01).backward()
        out.append((pos.detach(), vel.detach(), torch.stack(caller.grfs), torch.stack(caller.jafs),
                    [a[k].grad for k in KEYS]))
    # the kernel instances of the two layouts are separate compilations (different fma contraction): last-bit level
    # (joint forces amplify a last-bit pose difference by the 8e3..1.6e4 N/m attachment stiffness)
    gtol = 1e-4 if robot == "laikago" else 1e-5
    for other, layout in ((out[1], "latency"), (out[2], "team")):
        for name, tol, x, y in zip(("pos", "vel", "grf", "jaf"), (2e-6, 2e-6, 1e-4, 1e-4), out[0][:4], other[:4]):
            assert rel(x, y.double().cpu()) < tol, (layout, name, rel(x, y.double().cpu()))
        # gradients: 1e-5, except laikago in stiff contact where last-bit differences between two compilations are
        # amplified like every other rounding difference (fp32 noise floor of that fixture: 1e-3 .. 5e-3, tests/test_oracle.py)
        for k, x, y in zip(KEYS, out[0][4], other[4]):
            assert rel(x, y.double().cpu()) < gtol, (layout, k, rel(x, y.double().cpu()))


def test_single_frame_window_and_single_env():
    from oracle import sim_oracle as so
    from ppr_diffphys_b200 import SimEnv
    rm, d = make_inputs("quad", bs=1, T=1, seed=9)
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    a, _, _ = flat_args(d, dev)
    pos, vel, caller = run_cuda(env, a, 1, 1, 1)
    assert pos.shape == (1, rm.nb, 7) and len(caller.grfs) == 1
    bq, bqd = so.eval_fk(so.OracleModel(rm), d["q_init"].float().double(), d["qd_init"].float().double())
    assert (pos.cpu().double()[0] - bq[0]).abs().max() < 2e-6
    (pos.sum() + vel.sum()).backward()  # pure FK adjoint
    assert torch.isfinite(a["q_init"].grad).all() and float(a["refs"].grad.abs().max()) == 0.0


def test_c_abi_error_codes(monkeypatch):
    from ppr_diffphys_b200 import SimEnv, _lib
    monkeypatch.delenv("PPR_LATENCY_ENVS")
    lib = _lib.lib()
    env = SimEnv("laikago")
    h = env._h
    n = C.c_void_p(None)
    assert lib.ppr_fk_forward(h, 4, n, n, n, n, n) == -1                       # PPR_E_ARG
    assert lib.ppr_fk_forward(C.c_void_p(None), 4, n, n, n, n, n) == -3        # PPR_E_HANDLE
    assert lib.ppr_fk_forward(h, 0, n, n, n, n, n) == -1
    x = torch.zeros(1024, device="cuda")
    p = C.c_void_p(x.data_ptr())
    args = [p] * 2 + [n, n] + [p] * 6 + [p, p, n, n]
    assert lib.ppr_rollout_forward(h, 2, 65, 32, C.c_float(5e-4), 0, *args, p, C.c_size_t(16), n) == -4  # workspace
    assert lib.ppr_rollout_forward(h, 0, 65, 32, C.c_float(5e-4), 0, *args, p, C.c_size_t(16), n) == 0   # empty batch
    threads, epg = env.packing
    assert threads in (32, 96, 160) and epg == threads // env.nb
    assert env.latency_envs == 1024
    rowf = 24 if all(t in (1, 4) for t in env.model.joint_type) else 28    # floats per body and checkpoint row
    assert lib.ppr_rollout_workspace_bytes(h, 2, 65) == 2 * 65 * rowf * 32 * 4        # latency layout: 1 env per warp
    env.set_latency_envs(0)
    assert lib.ppr_rollout_workspace_bytes(h, 2, 65) == -(-2 // epg) * (threads // 32) * 65 * rowf * 32 * 4
    assert lib.ppr_model_set_latency_envs(h, -1) == -1
    assert env.team_envs == 296 and lib.ppr_model_set_team_envs(h, -1) == -1
    env.set_team_envs(5)
    assert env.team_envs == 5 and lib.ppr_model_team_envs(C.c_void_p(None)) == -3
    before = _lib.launch_count()
    env.fk(torch.zeros(3, env.nq, device="cuda"), torch.zeros(3, env.nqd, device="cuda"))
    assert _lib.launch_count() == before + 1
    # struct-argument entry points (ppr_rollout_io)
    from ppr_diffphys_b200._capi import RolloutIO
    assert lib.ppr_rollout_forward_ex(h, None, n) == -1
    io = RolloutIO()
    io.bs, io.nsteps, io.frame_stride, io.dt = 2, 65, 32, 5e-4
    assert lib.ppr_rollout_forward_ex(h, C.byref(io), n) == -1                         # missing pointers
    for f in ("q_init", "qd_init", "refs", "target_ke", "target_kd", "body_inv_mass", "body_inertia", "body_inv_inertia",
              "out_pos", "out_vel", "workspace"):
        setattr(io, f, x.data_ptr())
    io.workspace_bytes = 16
    assert lib.ppr_rollout_forward_ex(h, C.byref(io), n) == -4                         # workspace too small
    io.loss_pos = x.data_ptr()
    assert lib.ppr_rollout_forward_ex(h, C.byref(io), n) == -1                         # loss requested without targets
    io.loss_pos = None
    io.workspace_bytes = lib.ppr_rollout_workspace_bytes(h, 2, 65)
    assert lib.ppr_rollout_backward_ex(h, C.byref(io), n) == -1                        # no adjoint outputs / nothing to seed
    io.bs = 0
    assert lib.ppr_rollout_forward_ex(h, C.byref(io), n) == 0                          # empty batch
    assert lib.ppr_rollout_shared_grad_floats(h) == 2 * env.nqd + 19 * env.nb
    assert lib.ppr_refs_from_frames(65, 32, 0, 4, p, p, n) == -1 and lib.ppr_refs_from_frames(0, 32, 3, 4, p, p, n) == 0
    assert lib.ppr_model_set_ground(C.c_void_p(None), 1) == -3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_models_on_two_devices_in_one_process():
    """The > 48 kB dynamic shared-memory opt-in of the rollout kernels is per DEVICE, and the library must launch on the
    model's device whatever the caller's current device is: same rollout + adjoint on cuda:0 and cuda:1, bit-identical."""
    from ppr_diffphys_b200 import SimEnv
    stride, F, bs = 8, 3, 9
    T = stride * (F - 1) + 1
    rm, d = make_inputs("human", bs=bs, T=T, seed=7)
    d = settle_height(rm, d, 0.002)
    out = []
    for dev_id in (0, 1):
        dev = torch.device("cuda", dev_id)
        env = SimEnv(rm, device=dev)
        torch.cuda.set_device(0)                    # the caller's current device stays 0 for both
        a, _, _ = flat_args(d, dev)
        pos, vel, _ = run_cuda(env, a, bs, T, stride)
        ((pos ** 2).sum() + (vel ** 2).sum() * 0.01).backward()
        torch.cuda.synchronize(dev)
        out.append((pos.detach().cpu(), [a[k].grad.cpu() for k in KEYS]))
    assert torch.equal(out[0][0], out[1][0])
    for x, y in zip(out[0][1], out[1][1]):
        assert torch.equal(x, y)


def test_ground_flag_skips_the_contact_kernel():
    """``env.ground = False``: compute_forces skips eval_body_contacts and grf stays at res_f (integrator_euler.py:492-510)."""
    from ppr_diffphys_b200 import SimEnv
    stride, F, bs = 8, 3, 3
    T = stride * (F - 1) + 1
    rm, d = make_inputs("human", bs=bs, T=T, seed=2)
    d = settle_height(rm, d, 0.003)
    dev = torch.device("cuda:0")
    out = {}
    for ground in (True, False):
        env = SimEnv(rm)
        env.ground = ground
        a, _, _ = flat_args(d, dev, drop=("torques", "res_f"))
        pos, vel, caller = run_cuda(env, a, bs, T, stride)
        (pos.sum() + vel.sum()).backward()
        assert all(torch.isfinite(a[k].grad).all() for k in a if a[k] is not None and a[k].grad is not None)
        out[ground] = (pos.detach(), torch.stack(caller.grfs))
    assert float(out[True][1].abs().max()) > 1.0            # in contact
    assert float(out[False][1].abs().max()) == 0.0          # no ground: grf = res_f = 0
    assert float((out[True][0] - out[False][0]).abs().max()) > 1e-5


def test_joint_X_p_setter_changes_fk():
    from ppr_diffphys_b200 import SimEnv
    env = SimEnv("laikago")
    q = torch.zeros(1, env.nq, device="cuda")
    q[0, 6] = 1
    qd = torch.zeros(1, env.nqd, device="cuda")
    b0, _ = env.fk(q, qd)
    xp = env.joint_X_p.clone()
    xp[1, 0] += 0.1
    env.joint_X_p = xp
    b1, _ = env.fk(q, qd)
    assert abs(float(b1[0, 1, 0] - b0[0, 1, 0]) - 0.1) < 1e-6


@pytest.mark.parametrize("robot", ["laikago", "human"])
def test_per_env_joint_X_p_parity(robot):
    """lab4d assigns env.joint_X_p with ONE BLOCK PER ENV (dp_interface.py:454-465: per-instance bone lengths):
    FK, trajectories and gradients against the oracle given the same [bs, nb, 7] table; then back to the shared
    table."""
    from oracle import sim_oracle as so
    from ppr_diffphys_b200 import SimEnv
    stride, F, bs = 16, 3, 6
    T = stride * (F - 1) + 1
    rm, d = make_inputs(robot, bs=bs, T=T, seed=5)
    g = torch.Generator().manual_seed(9)
    xp = torch.as_tensor(rm.joint_X_p, dtype=torch.float64)[None].repeat(bs, 1, 1)
    xp[:, 1:, :3] *= 1.0 + 0.2 * (torch.rand(bs, 1, 1, generator=g, dtype=torch.float64) - 0.5)   # bone-length scale per env
    dq = torch.randn(bs, rm.nb, 4, generator=g, dtype=torch.float64) * 0.03
    q = xp[:, :, 3:] + dq
    xp[:, 1:, 3:] = (q / q.norm(dim=-1, keepdim=True))[:, 1:]
    xp = xp.float().double()
    m = so.OracleModel(rm)
    m.joint_X_p = xp
    # settle on the ground with the per-env skeleton
    bq, _ = so.eval_fk(m, d["q_init"], d["qd_init"])
    cb = torch.as_tensor(rm.contact_body, dtype=torch.long)
    pts = so.quat_rotate(bq[:, cb, 3:7], torch.as_tensor(rm.contact_point, dtype=torch.float64)[None].expand(bs, -1, 3))
    d["q_init"][:, 1] -= (bq[:, cb, 1] + pts[..., 1]).min(1)[0] + 0.002
    d = {k: v.float().double() for k, v in d.items()}
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    env.joint_X_p = xp.reshape(-1, 7).to(dev, torch.float32)
    # FK with T frames x bs envs: articulation i uses block i % bs
    q2 = torch.stack([d["q_init"], d["q_init"] * 0.9]).float().to(dev)          # [2, bs, nq]
    qd2 = torch.stack([d["qd_init"], d["qd_init"] * 0.5]).float().to(dev)
    fbq, fbqd = env.fk(q2.reshape(2 * bs, -1), qd2.reshape(2 * bs, -1))
    for t in range(2):
        obq, obqd = so.eval_fk(m, q2[t].double().cpu(), qd2[t].double().cpu())
        assert (fbq.view(2, bs, rm.nb, 7)[t].cpu().double() - obq).abs().max() < 5e-6
        assert (fbqd.view(2, bs, rm.nb, 6)[t].cpu().double() - obqd).abs().max() < 5e-5
    a, _, _ = flat_args(d, dev)
    pos, vel, caller = run_cuda(env, a, bs, T, stride)
    o = {k: d[k].clone().requires_grad_(True) for k in KEYS}
    opos, ovel, ogrf, ojaf = so.rollout(m, o["q_init"], o["qd_init"], o["torques"], o["res_f"], o["refs"],
                                        o["target_ke"], o["target_kd"], o["body_inv_mass"], o["body_inertia"],
                                        o["body_inv_inertia"], 5e-4, stride, F)
    assert ogrf.abs().max() > 1.0
    assert (pos.cpu().double() - opos.detach().reshape(F, -1, 7)).abs().max() <= POS_TOL
    wp = torch.randn(opos.shape, generator=g, dtype=torch.float64)
    wv = torch.randn(ovel.shape, generator=g, dtype=torch.float64) * 0.1
    torch.autograd.backward([pos, vel], [wp.reshape(F, -1, 7).to(dev, torch.float32),
                                         wv.reshape(F, -1, 6).to(dev, torch.float32)])
    grads = torch.autograd.grad((opos * wp).sum() + (ovel * wv).sum(), [o[k] for k in KEYS])
    tol = {k: GRAD_RTOL for k in KEYS}
    if robot == "laikago":
        tol, _ = grad_tolerances(rm, d, stride, F, adj_pos=wp, adj_vel=wv)   # (shared joint_X_p: same conditioning)
    for k, gr in zip(KEYS, grads):
        assert rel(a[k].grad, gr) <= tol[k], (k, rel(a[k].grad, gr), tol[k])
    # back to one shared table: same result as a fresh model
    env.joint_X_p = torch.as_tensor(rm.joint_X_p)
    a2, _, _ = flat_args(d, dev, requires_grad=False)
    p_shared, _, _ = run_cuda(env, a2, bs, T, stride)
    p_fresh, _, _ = run_cuda(SimEnv(rm), a2, bs, T, stride)
    assert torch.equal(p_shared, p_fresh) and not torch.equal(p_shared, pos.detach())


def test_forward_kinematics_reference_quirks():
    """dp_model.py:1030-1035,1080-1082: CPU inputs are moved to the GPU and the outputs back; :1109-1110 the
    gradient is scrubbed NaN -> 0 and clamped from ABOVE at +1 only."""
    from ppr_diffphys_b200 import ForwardKinematics, SimEnv
    env = SimEnv("laikago")
    T, bs = 2, 3
    rm, d = make_inputs("laikago", bs=T * bs, T=1, seed=4, ang=0.5)
    q = d["q_init"].float().view(T, bs, -1).clone().requires_grad_(True)       # CPU tensors
    qd = d["qd_init"].float().view(T, bs, -1).clone().requires_grad_(True)
    bq, bqd, frames = ForwardKinematics.apply(q, qd, env)
    assert not bq.is_cuda and bq.shape == (bs, T, rm.nb, 7)
    (bq[..., 0].sum() * 50.0 - bq[..., 1].sum() * 50.0).backward()             # +-50 per body on root x / y
    g = q.grad
    assert not g.is_cuda and torch.isfinite(g).all()
    assert float(g.max()) == 1.0                 # 13 bodies x 50 on root x -> clamped to +1
    assert float(g.min()) < -100.0               # no lower clamp (reference quirk)


def test_record_forces_flag_and_last_step_skip():
    from ppr_diffphys_b200 import SimEnv
    stride, F, bs = 8, 3, 4
    T = stride * (F - 1) + 1
    rm, d = make_inputs("human", bs=bs, T=T, seed=6)
    d = settle_height(rm, d, 0.003)
    dev = torch.device("cuda:0")
    env = SimEnv(rm)
    from ppr_diffphys_b200 import ForwardWarp
    outs = []
    for rec in (True, False):
        a, _, _ = flat_args(d, dev, requires_grad=False)
        caller = Caller(env, bs, T, stride)
        caller.record_forces = rec
        pos, vel = ForwardWarp.apply(a["q_init"], a["qd_init"], a["torques"], a["res_f"], a["refs"], a["target_ke"],
                                     a["target_kd"], a["body_mass"], a["body_inv_mass"], a["body_inertia"],
                                     a["body_inv_inertia"], caller)
        outs.append((pos, vel, caller))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert len(outs[0][2].grfs) == F and len(outs[1][2].grfs) == 0
    # grf = body_f after the contact pass: fn = c * ke < 0 is SUBTRACTED (integrator_euler.py:147,179), i.e. bodies in
    # contact carry an upward (+y) force
    grf = torch.stack(outs[0][2].grfs)
    assert float(grf[..., 4].max()) > 1.0 and float(grf[..., 4].min()) >= 0.0 and torch.isfinite(grf).all()
