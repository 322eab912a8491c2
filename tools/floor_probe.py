import os, sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from helpers import fp32_noise_floor, ROLLOUT_KEYS as KEYS, rel_err
from ppr_diffphys_b200 import load_robot
for fx in ["laikago", "human", "quad", "laikago_air"]:
    z = np.load("tests/golden/rollout_%s.npz" % fx)
    rm = load_robot(str(z["robot"]))
    d = {k: torch.from_numpy(z["in_" + k]) for k in KEYS}
    stride, F = int(z["stride"]), int(z["nframes"])
    floor, g64 = fp32_noise_floor(rm, d, stride, F, adj_pos=torch.from_numpy(z["adj_pos"]), adj_vel=torch.from_numpy(z["adj_vel"]))
    gold = {k: rel_err(g64[k], torch.from_numpy(z["grad_" + k])) for k in KEYS}
    print(fx, "floor", {k: "%.1e" % v for k, v in floor.items()})
    print(fx, "oracle64 vs golden max", max(gold.values()))
    if torch.cuda.is_available():
        from test_gpu_parity import flat_args, run_cuda
        from ppr_diffphys_b200 import SimEnv
        for lat in ("0", "1000000"):
            os.environ["PPR_LATENCY_ENVS"] = lat
            env = SimEnv(rm)
            a, bs, T = flat_args(d, torch.device("cuda:0"))
            pos, vel, _ = run_cuda(env, a, bs, T, stride)
            dev = pos.device
            torch.autograd.backward([pos, vel], [torch.from_numpy(z["adj_pos"]).reshape(F, -1, 7).to(dev, torch.float32), torch.from_numpy(z["adj_vel"]).reshape(F, -1, 6).to(dev, torch.float32)])
            print(fx, "cuda lat=%s" % lat, {k: "%.1e" % rel_err(a[k].grad, g64[k]) for k in KEYS})
