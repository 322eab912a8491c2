"""Small fwd+bwd rollouts of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
four robots x {throughput layout, latency layout, throughput layout + checkpoint-every-3 recompute + per-env
joint_X_p}.  usage: compute-sanitizer --tool <tool> python tools/sanitize_smoke.py"""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from helpers import make_inputs, settle_height, make_mixed_robot
from test_gpu_parity import flat_args, run_cuda
from ppr_diffphys_b200 import SimEnv
for robot in ['laikago', 'human', 'quad', make_mixed_robot()]:
    stride, F, bs = 4, 3, 9
    T = stride * (F - 1) + 1
    rm, d = make_inputs(robot, bs=bs, T=T, seed=3, res_f_std=0.05, torque_std=0.05)
    d = settle_height(rm, d, 0.003)
    for mode in ('throughput', 'latency', 'recompute+per-env-X_p'):
        env = SimEnv(rm)
        env.set_latency_envs(1 << 20 if mode == 'latency' else 0)
        if mode.startswith('recompute'):
            env.set_checkpoint_every(3)
            env.joint_X_p = torch.as_tensor(rm.joint_X_p).repeat(bs, 1).cuda()
        a, _, _ = flat_args(d, torch.device('cuda:0'))
        pos, vel, _ = run_cuda(env, a, bs, T, stride)
        (pos.sum() + vel.sum()).backward()
        torch.cuda.synchronize()
        print(rm.name, mode, env.packing, 'ok', float(pos.abs().max()))
