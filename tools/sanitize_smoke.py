"""Small fwd+bwd rollouts of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
four robots x {throughput layout, latency layout, team layout, throughput layout + checkpoint-every-3 recompute +
per-env joint_X_p}, plus the round-2 paths: shared (un-replicated) parameters with the epilogue reduction, the fused pose
loss through the struct-argument entry points, and the refs-from-frames kernels.
usage: compute-sanitizer --tool <tool> python tools/sanitize_smoke.py [mode substring, e.g. team]"""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helpers import make_inputs, settle_height, make_mixed_robot
from test_gpu_parity import flat_args, run_cuda
from ppr_diffphys_b200 import ForwardWarp, ForwardWarpLoss, RefsFromFrames, SimEnv
only = sys.argv[1] if len(sys.argv) > 1 else ''
for robot in ['laikago', 'human', 'quad', make_mixed_robot()]:
    stride, F, bs = 4, 3, 9
    T = stride * (F - 1) + 1
    rm, d = make_inputs(robot, bs=bs, T=T, seed=3, res_f_std=0.05, torque_std=0.05)
    d = settle_height(rm, d, 0.003)
    for mode in ('throughput', 'latency', 'team', 'recompute+per-env-X_p'):
        if only not in mode:
            continue
        env = SimEnv(rm)
        env.set_latency_envs(1 << 20 if mode in ('latency', 'team') else 0)
        env.set_team_envs(1 << 20 if mode == 'team' else 0)
        if mode.startswith('recompute'):
            env.set_checkpoint_every(3)
            env.joint_X_p = torch.as_tensor(rm.joint_X_p).repeat(bs, 1).cuda()
        a, _, _ = flat_args(d, torch.device('cuda:0'))
        pos, vel, _ = run_cuda(env, a, bs, T, stride)
        (pos.sum() + vel.sum()).backward()
        torch.cuda.synchronize()
        print(rm.name, mode, env.packing, 'ok', float(pos.abs().max()))


class _Caller:
    def __init__(self, env, n, T, stride):
        self.env, self.num_envs, self.dt = env, n, 5e-4
        self.steps_idx, self.frame2step, self.record_forces = range(T), [i for i in range(T) if i % stride == 0], False


for robot in ['laikago', 'human']:
    stride, F, bs = 4, 3, 9
    T = stride * (F - 1) + 1
    rm, d = make_inputs(robot, bs=bs, T=T, seed=3)
    d = settle_height(rm, d, 0.003)
    dev = torch.device('cuda:0')
    for lat, team in ((0, 0), (1 << 20, 0), (1 << 20, 1 << 20)):
        if only and not (only == 'team' and team):
            continue
        env = SimEnv(rm)
        env.set_latency_envs(lat)
        env.set_team_envs(team)
        t = lambda x: torch.as_tensor(x, dtype=torch.float32, device=dev)
        m, nI = t(rm.body_mass), t(rm.norm_body_inertia)
        leaf = lambda x: x.clone().requires_grad_(True)
        frames = leaf(d['refs'].float().reshape(T, -1)[::stride].contiguous().to(dev))
        refs = RefsFromFrames.apply(frames, stride, T)
        p = [leaf(x) for x in (t(rm.joint_target_ke), t(rm.joint_target_kd), 1 / m, nI * m[:, None, None],
                               torch.linalg.inv(nI * m[:, None, None]))]
        q, qd = leaf(d['q_init'].float().reshape(-1).to(dev)), leaf(d['qd_init'].float().reshape(-1).to(dev))
        pos, vel = ForwardWarp.apply(q, qd, None, None, refs, p[0], p[1], m, p[2], p[3], p[4], _Caller(env, bs, T, stride))
        (pos.sum() + vel.sum()).backward()
        tgt = (pos.detach() + 0.01).requires_grad_(True)
        loss, pos2, vel2 = ForwardWarpLoss.apply(q, qd, None, None, refs.detach(), p[0], p[1], m, p[2], p[3], p[4], tgt, 0.1,
                                                 _Caller(env, bs, T, stride))
        loss.sum().backward()
        torch.cuda.synchronize()
        print(rm.name, 'shared + fused loss + refs-from-frames', env.packing, 'lat', lat, 'team', team, 'ok', float(loss.sum()))
