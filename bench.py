#!/usr/bin/env python
"""bench.py -- env-steps/s of the rollout hot path (forward + hand-written adjoint) on N B200s.

    python bench.py --gpus 1 --steps K --warmup W                      # single GPU
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W     # CPU arm (rank 0 only)

A "step" = one optimisation step of the hot path on one batch of synthetic input: ForwardWarp forward
(one persistent rollout kernel), a quadratic pose loss, ForwardWarp backward (one adjoint kernel), torch autograd
chaining the four mass-related gradients into body_mass (dp_model.py:725-730), and -- for N > 1 -- the single
NCCL all-reduce of the packed shared-parameter gradients (target_ke, target_kd, body_mass).  1 env-step = one
simulation substep of one environment.

Workload (BASELINE.json configs[4], the scaling sweep, at its largest per-GPU size): laikago, 65 536 envs per GPU
(weak scaling), 64 differentiated substeps per window, synthetic inputs (seed 0 + rank).  Other configs are
parity-test cases; `--workload` selects them for ad-hoc measurements and `--extras` appends their numbers.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: robot, per-GPU envs, differentiated substeps, frame stride, clearance (m; <0 = penetrating), lin_vel
    "laikago-scaling-65536x64": dict(robot="laikago", bs=65536, window=64, stride=32, clearance=1e-3, lin_vel=0.0),
    "laikago-trot-64x760": dict(robot="laikago", bs=64, window=759, stride=33, clearance=1e-3, lin_vel=0.0),
    "quad-1024x64": dict(robot="quad", bs=1024, window=64, stride=32, clearance=1e-3, lin_vel=0.0),
    "human-4096x64-contact": dict(robot="human", bs=4096, window=64, stride=32, clearance=-0.0025, lin_vel=1.0),
    "human-65536x64-contact": dict(robot="human", bs=65536, window=64, stride=32, clearance=-0.0025, lin_vel=1.0),
}
DEFAULT_WORKLOAD = "laikago-scaling-65536x64"
DT = 5e-4


def algorithmic_bytes(nb, nqd):
    """S = nb*13*4 (body_q + body_qd), R = nqd*4. fwd = 2S+R, bwd = 3S+2R, fwd+bwd = 5S+3R."""
    S, R = nb * 13 * 4, nqd * 4
    return dict(fwd=2 * S + R, bwd=3 * S + 2 * R, fwdbwd=5 * S + 3 * R)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_measure(robot, window, stride, target_seconds=12.0, steps=1, warmup=0):
    """Times the CPU port (oracle/cpu_port, OpenMP over envs, all host threads) on a bounded sample of the
    workload: bs_sample envs x `window` substeps, forward + reverse sweep. Returns env-steps/s."""
    import torch
    from oracle.cpu_port import CpuRollout, use_all_cores
    from ppr_diffphys_b200 import load_robot
    rm = load_robot(robot)
    cores = use_all_cores()
    nsteps = window + 1
    F = (nsteps - 1) // stride + 1

    def make(bs):
        g = torch.Generator().manual_seed(0)
        B = rm.nqd - 6
        ja = (torch.rand(bs, B, generator=g) * 2 - 1) * 0.2
        q = torch.zeros(bs, rm.nq); q[:, 6] = 1.0; q[:, 7:] = ja
        q[:, 1] = 0.45
        refs = torch.zeros(nsteps, bs, rm.nqd); refs[:, :, 6:] = ja[None]
        mass = torch.as_tensor(rm.body_mass)[None].repeat(bs, 1)
        I = torch.as_tensor(rm.norm_body_inertia)[None] * mass[..., None, None]
        return dict(q_init=q, qd_init=torch.randn(bs, rm.nqd, generator=g) * 0.1, torques=None, res_f=None, refs=refs,
                    target_ke=torch.as_tensor(rm.joint_target_ke)[None].repeat(bs, 1).contiguous(),
                    target_kd=torch.as_tensor(rm.joint_target_kd)[None].repeat(bs, 1).contiguous(),
                    body_inv_mass=1.0 / mass, body_inertia=I.contiguous(), body_inv_inertia=torch.linalg.inv(I))

    cpu = CpuRollout(rm)

    def one(d):
        t0 = time.perf_counter()
        pos, vel = cpu.forward(d, DT, stride, F)
        cpu.backward(pos * 0.01, vel * 0.01)
        return time.perf_counter() - t0

    probe_bs = 4 * cores
    t_probe = one(make(probe_bs))
    bs = int(max(cores, min(65536, probe_bs * target_seconds / max(t_probe, 1e-6) / max(1, steps + warmup))))
    bs = max(cores, (bs // cores) * cores)
    d = make(bs)
    for _ in range(warmup):
        one(d)
    times = [one(d) for _ in range(max(1, steps))]
    t = sum(times) / len(times)
    return dict(value=bs * window / t, cores=cores, sample="%s, %d envs x %d substeps fwd+bwd, fp32 CPU port "
                "(all %d contact points/env tested per substep like the reference), %d timed step(s), %.2f s/step"
                % (robot, bs, window, rm.nc, len(times), t), ms_per_step=t * 1e3, bs=bs)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    r = cpu_measure(w["robot"], w["window"], w["stride"], target_seconds=20.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "env_steps_per_sec_fwd_bwd", "value": r["value"], "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "robot": w["robot"], "substeps_per_window": w["window"],
                   "note": "reference's Warp CPU device cannot be installed here (no network); this is the repo's "
                           "C++/OpenMP port of the reference kernels on the box's host cores, bounded sample"},
        "cpu_baseline": {"value": r["value"], "unit": "env-steps/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class Caller:
    def __init__(self, env, num_envs, nsteps, stride):
        self.env, self.num_envs, self.dt = env, num_envs, DT
        self.steps_idx = range(nsteps)
        self.frame2step = [i for i in range(nsteps) if i % stride == 0]
        self.record_forces = False


def kernel_source_hash():
    """sha256 over the CUDA sources: profiles/traffic.json is stamped with it by tools/update_traffic.py, and a stale
    stamp (the kernels changed since the ncu capture) is reported as such instead of being passed off as current."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "ppr_diffphys_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".h")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def step_functions():
    """The caller-side glue of the bench step as two autograd Functions with hand-written backwards: the same values and
    gradients as the plain torch expressions in their docstrings (tests/test_bench_host.py compares them), in a third of
    the device ops -- at the small configs the step is launch-bound and every elementwise op is a graph node."""
    import torch

    class StepLoss(torch.autograd.Function):
        """(pos[-1,:,:3] - pos[0,:,:3]).pow(2).mean() + 1e-3 * vel[-1].pow(2).mean()"""

        @staticmethod
        def forward(ctx, pos, vel):
            d = pos[-1, :, :3] - pos[0, :, :3]
            v = vel[-1]
            ctx.save_for_backward(d, v)
            ctx.shapes = (pos.shape, vel.shape)
            return torch.add(torch.dot(d.reshape(-1), d.reshape(-1)) / d.numel(),
                             torch.dot(v.reshape(-1), v.reshape(-1)), alpha=1e-3 / v.numel())

        @staticmethod
        def backward(ctx, g):
            d, v = ctx.saved_tensors
            gpos = torch.zeros(ctx.shapes[0], device=d.device, dtype=d.dtype)
            gvel = torch.zeros(ctx.shapes[1], device=d.device, dtype=d.dtype)
            gd = d * (g * (2.0 / d.numel()))
            gpos[-1, :, :3] = gd
            torch.neg(gd, out=gpos[0, :, :3]) if gpos.shape[0] > 1 else gpos[0, :, :3].zero_()
            torch.mul(v, g * (2e-3 / v.numel()), out=gvel[-1])
            return gpos, gvel

    class MassChain(torch.autograd.Function):
        """(1 / m, nI * m[:, None, None], nI_inv / m[:, None, None]): what dp_model.py:726-730 derives from body_mass"""

        @staticmethod
        def forward(ctx, m, nI, nI_inv):
            inv_m = 1.0 / m
            ctx.save_for_backward(inv_m, nI, nI_inv)
            return inv_m, nI * m[:, None, None], nI_inv * inv_m[:, None, None]

        @staticmethod
        def backward(ctx, g_inv_m, g_I, g_inv_I):
            inv_m, nI, nI_inv = ctx.saved_tensors
            a = (g_inv_I * nI_inv).sum((1, 2)).add_(g_inv_m)         # d/d(1/m)
            return torch.addcmul((g_I * nI).sum((1, 2)), a, inv_m * inv_m, value=-1.0), None, None

    return StepLoss, MassChain


def workload_batch(env, w, bs, nsteps, seed, pinned_host=False):
    """The synthetic inputs of a bench workload -- also what tests/test_gpu_bench_parity.py checks against the oracle."""
    from ppr_diffphys_b200.synth import make_batch
    return make_batch(env, bs, nsteps, seed=seed, clearance=w["clearance"], lin_vel=w["lin_vel"],
                      pinned_host=pinned_host, frame_stride=w["stride"])


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from ppr_diffphys_b200 import ForwardWarp, RefsFromFrames, SimEnv, _lib, load_robot
    from ppr_diffphys_b200.synth import shared_param_chain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # everything runs on an ordinary (non-legacy) stream: autograd binds the gradient-accumulation nodes of the leaves
    # to the stream of their first backward, and the legacy default stream cannot take part in a CUDA-graph capture
    torch.cuda.set_stream(torch.cuda.Stream())
    w = WORKLOADS[args.workload]
    if args.scaling == "strong":
        total = args.total_envs if args.total_envs else w["bs"]
        bs = total // world                       # fixed total work, sharded
    else:
        bs = args.envs if args.envs else w["bs"]  # fixed work per GPU
    window, stride = w["window"], w["stride"]
    nsteps = window + 1
    rm = load_robot(w["robot"])
    env = SimEnv(rm)
    if args.ckpt_every > 1:
        env.set_checkpoint_every(args.ckpt_every)
    nb, nqd = rm.nb, rm.nqd
    caller = Caller(env, bs, nsteps, stride)
    host = workload_batch(env, w, bs, nsteps, seed=rank, pinned_host=True)
    nI = torch.as_tensor(rm.norm_body_inertia, device=dev)
    nI_inv = torch.linalg.inv(nI)
    # shared parameters (what the reference optimises): PD gains + body mass
    p_ke = torch.as_tensor(rm.joint_target_ke, device=dev).clone().requires_grad_(True)
    p_kd = torch.as_tensor(rm.joint_target_kd, device=dev).clone().requires_grad_(True)
    p_mass = torch.as_tensor(rm.body_mass, device=dev).clone().requires_grad_(True)
    NP = 2 * nqd + nb + 1                         # shared-parameter gradients + the loss: one D2H read / one all-reduce
    StepLoss, MassChain = step_functions()

    def step(inp, replicate=False, zero_forces=False):
        """one optimisation step of the hot path on device-resident inputs -> packed [grad ke | grad kd | grad mass | loss]"""
        q_init = inp["q_init"].detach().requires_grad_(True)
        qd_init = inp["qd_init"].detach().requires_grad_(True)
        refs = inp["refs"].detach().requires_grad_(True)
        if replicate:   # the reference's literal calling convention: per-env replicated parameters (dp_model.py:723-730)
            ke, kd, mass, inv_m, I, inv_I = shared_param_chain(p_ke, p_kd, p_mass, nI, bs)
        else:           # un-replicated parameters: the kernels read one shared copy, gradients are reduced on the device
            ke, kd, mass = p_ke, p_kd, p_mass
            inv_m, I, inv_I = MassChain.apply(p_mass, nI, nI_inv)   # inverse(nI * m) = inverse(nI) / m
        torques = res_f = None
        if zero_forces:  # exact zeros passed as real tensors, as the reference does (dp_model.py:529,536)
            torques, res_f = inp["torques0"], inp["res_f0"]
        pos, vel = ForwardWarp.apply(q_init, qd_init, torques, res_f, refs, ke, kd, mass, inv_m, I, inv_I, caller)
        loss = StepLoss.apply(pos, vel)
        # every gradient the reference's backward returns: shared parameters (.grad of the three leaves) + per-env control
        # references / initial state (.grad of the per-step leaves)
        for p in (p_ke, p_kd, p_mass):
            p.grad = None
        loss.backward()
        return torch.cat([p_ke.grad, p_kd.grad, p_mass.grad, loss.detach().reshape(1)])

    keys = ("q_init", "qd_init", "refs")
    static = {k: host[k].to(dev) for k in keys}            # inputs of the (graph-captured) step, resident in HBM
    bytes_full = sum(host[k].numel() * 4 for k in keys)
    ckeys = ("q_init", "qd_init", "ref_frames")
    bytes_compact = sum(host[k].numel() * 4 for k in ckeys)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # ---- warm-up (eager), then capture the step ONCE in a CUDA graph: fixed shapes, so every later step is one replay
    # (the small / medium configs are launch-bound: 30-odd launches + autograd bookkeeping per step)
    for _ in range(max(3, args.warmup)):
        step(static)
    graph, packed = None, None
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step(static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n_before = _lib.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            packed = step(static)
        launches_per_step = _lib.launch_count() - n_before

        def run_step():
            graph.replay()
            if world > 1:
                dist.all_reduce(packed)
            return packed
    else:
        n_before = _lib.launch_count()
        step(static)
        launches_per_step = _lib.launch_count() - n_before

        def run_step():
            out = step(static)
            if world > 1:
                dist.all_reduce(out)
            return out
    for _ in range(3):
        run_step()
    # ---- device-resident timing (value)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(run_step, args.steps)
    # the same step launched eagerly (no graph), for the host-overhead comparison
    def eager_step():
        out = step(static)
        if world > 1:
            dist.all_reduce(out)
    for _ in range(2):   # (the first eager steps after the capture re-grow the allocator's ordinary pool)
        eager_step()
    ms_eager = timed(eager_step, max(3, args.steps // 4)) / max(3, args.steps // 4)
    if args.census and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(static)
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda r: -r.count)
        print("kernel census of one step: %d device ops" % sum(r.count for r in rows), file=sys.stderr)
        for r in rows:
            print("  %3d x %8.1f us  %s" % (r.count, r.device_time_total / max(r.count, 1), r.key[:110]), file=sys.stderr)
    # ---- kernel-only timing with CUDA events on the launching stream (roofline)
    ke = p_ke.detach()[None].expand(bs, nqd).reshape(-1).contiguous()
    kd = p_kd.detach()[None].expand(bs, nqd).reshape(-1).contiguous()
    mass = p_mass.detach()[None].expand(bs, nb).reshape(-1).contiguous()
    inv_m, I = 1.0 / mass, (nI[None].expand(bs, nb, 3, 3).reshape(-1, 3, 3) * mass[:, None, None]).contiguous()
    inv_I = (nI_inv[None].expand(bs, nb, 3, 3).reshape(-1, 3, 3) * inv_m[:, None, None]).contiguous()
    ws = None
    evs = []
    for i in range(max(args.steps, 5) + 1):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        pos, vel, _, _, ws = env.rollout_forward(bs, nsteps, stride, DT, static["q_init"], static["qd_init"], None,
                                                 None, static["refs"], ke, kd, inv_m, I, inv_I, want_forces=False,
                                                 workspace=ws)
        e[1].record()
        env.rollout_backward(bs, nsteps, stride, DT, static["q_init"], static["qd_init"], None, None,
                             static["refs"], ke, kd, inv_m, I, inv_I, pos, vel, ws)
        e[2].record()
        evs.append(e)
    torch.cuda.synchronize()
    del ws, pos, vel, ke, kd, mass, inv_m, I, inv_I
    med = lambda v: sorted(v)[len(v) // 2]   # median: robust to a stray slow launch
    fwd_ms = med([e[0].elapsed_time(e[1]) for e in evs[1:]])
    bwd_ms = med([e[1].elapsed_time(e[2]) for e in evs[1:]])

    # ---- end to end through the public call: what a caller SHIPS per step is q_init, qd_init and the per-FRAME control
    # references (pinned host memory -> device, on a copy stream, double buffered so that the copy of step i+1 overlaps
    # the kernels of step i); the per-substep references are expanded on the device (RefsFromFrames: the reference
    # interpolates its mocap targets per substep on the host, dp_model.py:421-427,605-609); loss + packed shared
    # gradients come back to pinned host memory every step.
    copy_stream = torch.cuda.Stream()

    def make_e2e(ship_keys):
        bufs = [{k: torch.empty_like(host[k], device=dev) for k in ship_keys} for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        host_out = [torch.empty(NP).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]

        def enqueue_copy(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[i % 2])
                for k in ship_keys:
                    bufs[i % 2][k].copy_(host[k], non_blocking=True)
                ready[i % 2].record(copy_stream)

        def run(steps):
            cur = torch.cuda.current_stream()
            for b in range(2):
                free[b].record(cur)
            enqueue_copy(0)
            out = None
            for i in range(steps):
                if i + 1 < steps:
                    enqueue_copy(i + 1)
                cur.wait_event(ready[i % 2])
                b = bufs[i % 2]
                static["q_init"].copy_(b["q_init"], non_blocking=True)
                static["qd_init"].copy_(b["qd_init"], non_blocking=True)
                if "ref_frames" in b:    # device-side expansion of the shipped per-frame references
                    _lib.check(_lib.lib().ppr_refs_from_frames(nsteps, stride, b["ref_frames"].shape[0],
                                                               b["ref_frames"].shape[1], b["ref_frames"].data_ptr(),
                                                               static["refs"].data_ptr(), cur.cuda_stream),
                               "ppr_refs_from_frames")
                else:
                    static["refs"].copy_(b["refs"], non_blocking=True)
                free[i % 2].record(cur)
                res = run_step()
                host_out[i % 2].copy_(res, non_blocking=True)   # device -> host every step
                done[i % 2].record(cur)
                if i > 0:
                    done[(i - 1) % 2].synchronize()
                    out = float(host_out[(i - 1) % 2][-1])      # the host reads step i-1's loss
            done[(steps - 1) % 2].synchronize()
            return float(host_out[(steps - 1) % 2][-1])
        return run

    e2e_compact = make_e2e(ckeys)
    e2e_compact(2)
    ms_e2e = timed(lambda: e2e_compact(args.steps), 1)
    e2e_full = make_e2e(keys)
    e2e_full(2)
    ms_e2e_full = timed(lambda: e2e_full(args.steps), 1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the reference's literal calling convention: per-env replicated parameters AND exact-zero torques /
    # res_f passed as real tensors (their gradients are computed and returned), eager
    ref_conv = None
    if not args.no_extras or args.reference_convention:
        try:
            conv = dict(static, torques0=torch.zeros(nsteps, bs * nqd, device=dev, requires_grad=True),
                        res_f0=torch.zeros(nsteps, bs * nb, 6, device=dev, requires_grad=True))
            for _ in range(2):
                step(conv, replicate=True, zero_forces=True)
            k = max(3, args.steps // 4)
            ms_conv = timed(lambda: step(conv, replicate=True, zero_forces=True), k) / k
            ref_conv = {"value": bs * window * world / (ms_conv * 1e-3), "unit": "env-steps/s", "ms_per_step": ms_conv,
                        "note": "per-env replicated ke/kd/mass/inertia + real zero torques / res_f tensors "
                                "(dp_model.py:529,536,723-730), eager, no all-reduce"}
            del conv
        except Exception as ex:
            ref_conv = {"error": repr(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    env_steps = bs * window * world
    ab = algorithmic_bytes(nb, nqd)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    # DRAM traffic of the dominant kernel per launch: from the committed ncu capture IF it was taken with these sources
    traffic, traffic_note, fp32 = None, None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj.get("kernel_source_hash") != kernel_source_hash():
            traffic_note = "profiles/traffic.json is stale (kernel sources changed since the ncu capture %s)" % tj.get(
                "kernel_source_hash")
        elif bs == w["bs"] and args.workload in tj:
            traffic = tj[args.workload]["rollout_backward_kernel"]
            fl = tj[args.workload].get("fp32_flops_per_env_step")
            if fl:   # secondary (honest) compute bound, SURVEY.md 8(d)
                mhz = (clocks or {}).get("sm_mhz") or 1965.0
                peak_tf = 148 * 128 * 2 * mhz * 1e6 / 1e12
                f_tf = fl["rollout_forward_kernel"] * bs * window / (fwd_ms * 1e-3) / 1e12
                b_tf = fl["rollout_backward_kernel"] * bs * window / (bwd_ms * 1e-3) / 1e12
                fp32 = {"flops_per_env_step": fl, "peak_tflops": peak_tf, "forward_tflops": f_tf, "backward_tflops": b_tf,
                        "forward_frac": f_tf / peak_tf, "backward_frac": b_tf / peak_tf}
    except Exception as ex:
        traffic_note = "profiles/traffic.json unreadable: %r" % (ex,)
    bwd_gbs = ab["bwd"] * bs * window / (bwd_ms * 1e-3) / 1e9
    fwd_gbs = ab["fwd"] * bs * window / (fwd_ms * 1e-3) / 1e9
    value = env_steps / (ms / args.steps * 1e-3)
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            r = cpu_measure(w["robot"], window, stride, target_seconds=10.0)
            cpu = {"value": r["value"], "unit": "env-steps/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        except Exception as ex:  # the checker is optional for the GPU arm
            cpu = {"value": None, "unit": "env-steps/s", "cores": None, "kind": "port", "sample": "failed: %r" % (ex,)}
    others = None
    if world == 1 and not args.no_extras and args.workload == DEFAULT_WORKLOAD and not args.envs and args.scaling == "weak":
        # the remaining BASELINE.json configs (parity-test cases, not bench lines) + the humanoid headline size, for
        # context: each in its own short process
        others = []
        for wname, st in (("human-65536x64-contact", 10), ("laikago-trot-64x760", 10), ("quad-1024x64", 20),
                          ("human-4096x64-contact", 20)):
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", wname, "--no-cpu",
                                      "--no-extras", "--steps", str(st), "--warmup", "3"], capture_output=True, text=True,
                                     timeout=300).stdout.strip().splitlines()[-1]
                d = json.loads(out)
                others.append({"workload": wname, "value": d["value"], "unit": d["unit"], "steps": d["steps"],
                               "ms_per_step": d["ms_per_step"], "ms_per_step_eager": d["ms_per_step_eager"],
                               "kernels_ms": d["kernels_ms"],
                               "step_over_kernels": d["ms_per_step"] / (d["kernels_ms"]["rollout_forward"] + d["kernels_ms"]["rollout_backward"]),
                               "roofline_frac_fwd_bwd": d["roofline"]["fwd_bwd_combined_frac"],
                               "fwd_only": d["fwd_only"]["value"], "e2e": d["e2e"]["value"]})
            except Exception as ex:
                others.append({"workload": wname, "error": repr(ex)})
        # BASELINE.json configs[0] (the reference's own recipe: laikago, mi-pace clip, 10 windows x 24 frames = 760
        # substeps, 3 MLPs + FK + rollout + se3 losses + backward + AdamW): seconds per optimisation iteration through
        # the caller loop of ppr_diffphys_b200.imitation, one CUDA-graph replay per iteration
        try:
            ex_py = os.path.join(ROOT, "examples", "run_imitation.py")
            out = subprocess.run([sys.executable, ex_py, "--seqname", "mi-pace", "--iters", "41", "--log-every", "1000"],
                                 capture_output=True, text=True, timeout=300).stdout.strip().splitlines()[-1]
            d = json.loads(out)
            others.append({"workload": "imitation-mi-pace-10x760", "ms_per_iteration": d["median_iter_ms"],
                           "value": d["env_steps_per_sec"], "unit": "env-steps/s", "cuda_graph": d["cuda_graph"],
                           "loss_traj_first": d["loss_traj_first"], "loss_traj_last": d["loss_traj_last"]})
        except Exception as ex:
            others.append({"workload": "imitation-mi-pace-10x760", "error": repr(ex)})
    line = {
        "metric": "env_steps_per_sec_fwd_bwd", "value": value, "unit": "env-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "robot": w["robot"], "envs_per_gpu": bs, "total_envs": bs * world,
                   "substeps_per_window": window, "frame_stride": stride, "bodies": nb, "dofs": nqd,
                   "contacts_per_env": rm.nc,
                   "refs": "per-frame samples of ja + 0.1 sin(2 pi t / 64 + phi), linear in between (mocap-style)",
                   "parallelism": "env-sharded x%d, 1 all-reduce of %d floats/step" % (world, NP),
                   "params": "shared (un-replicated), gradients reduced over envs in the adjoint kernel's epilogue",
                   "step": "one CUDA-graph replay (forward kernel, loss, adjoint kernel, reduce, autograd chain)"
                           if graph is not None else "eager launches",
                   "math": "-prec-div=false -prec-sqrt=false -ftz=true (MUFU rcp / sqrt, <= 2 ulp; Warp's default is IEEE)",
                   "checkpoint_every": args.ckpt_every,
                   "l2": "working set >> 126 MB L2 (state checkpoint %.2f GB/step streamed once each way)"
                         % (env.workspace_bytes(bs, nsteps) / 1e9)},
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "ms_per_step_eager": ms_eager,
        "kernels_ms": {"rollout_forward": fwd_ms, "rollout_backward": bwd_ms},
        "fwd_only": {"value": bs * window * world / (fwd_ms * 1e-3), "unit": "env-steps/s", "ms": fwd_ms,
                     "note": "forward rollout kernel alone (CUDA events), per GPU x n_gpus"},
        "roofline": {"bound": "hbm", "kernel": "rollout_backward_kernel", "achieved": bwd_gbs, "peak": peak_gbs,
                     "unit": "GB/s", "frac": bwd_gbs / peak_gbs, "traffic": traffic, "traffic_note": traffic_note,
                     "peak_source": peak_src, "algorithmic_bytes_per_env_step": ab,
                     "forward_kernel": {"achieved": fwd_gbs, "frac": fwd_gbs / peak_gbs},
                     "fwd_bwd_combined_frac": ab["fwdbwd"] * bs * window / ((fwd_ms + bwd_ms) * 1e-3) / 1e9 / peak_gbs,
                     "step_frac": ab["fwdbwd"] * bs * window / (ms / args.steps * 1e-3) / 1e9 / peak_gbs,
                     "fp32": fp32,
                     "note": "the path is latency / issue bound, not HBM bound (SURVEY.md 8d, profiles/README.md)"},
        "e2e": {"value": env_steps / (ms_e2e / args.steps * 1e-3), "unit": "env-steps/s",
                "h2d_bytes_per_step": bytes_compact, "d2h_bytes_per_step": 4 * NP, "ms_per_step": ms_e2e / args.steps,
                "note": "ships q_init, qd_init and per-FRAME control references from pinned host memory (copy of step "
                        "i+1 overlaps the kernels of step i), expands the per-substep references on the device"},
        "e2e_full_refs_from_host": {"value": env_steps / (ms_e2e_full / args.steps * 1e-3), "unit": "env-steps/s",
                                    "h2d_bytes_per_step": bytes_full, "ms_per_step": ms_e2e_full / args.steps,
                                    "note": "round-1 definition: every per-substep reference shipped from the host"},
        "reference_convention": ref_conv,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "other_configs": others,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--ckpt-every", type=int, default=1,
                    help="checkpoint policy K: keep the state every K substeps, recompute the rest in the adjoint")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the other_configs context measurements")
    ap.add_argument("--reference-convention", action="store_true",
                    help="also time the reference's literal convention (replicated parameters, real zero torques/res_f) "
                         "even with --no-extras")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of one CUDA-graph replay")
    ap.add_argument("--census", action="store_true",
                    help="print the kernel census of one step (torch.profiler, outside every timed region) to stderr")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --envs per GPU (default); strong: --total-envs sharded over the ranks")
    ap.add_argument("--total-envs", type=int, default=0, help="total envs of a strong-scaling run (default: workload size)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
