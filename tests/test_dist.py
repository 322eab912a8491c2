"""world_size-2 gloo tests of the multi-GPU host logic: contiguous env sharding with no data-path collective,
and the single all-reduce of the packed shared-parameter gradients.  The per-rank 'simulator' is the CPU port
(checker infrastructure) so the sharded-vs-unsharded gradient equality is exercised end to end on CPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import make_inputs
        from oracle.cpu_port import CpuRollout
        from ppr_diffphys_b200.dist import PackedGrads, shard_envs, shard_range
        torch.set_num_threads(1)
        bs, stride, F = 5, 4, 2
        T = stride * (F - 1) + 1
        rm, d = make_inputs("laikago", bs=bs, T=T, seed=3, height=1.0)
        ke = torch.nn.Parameter(torch.as_tensor(rm.joint_target_ke, dtype=torch.float64))
        mass = torch.nn.Parameter(torch.as_tensor(rm.body_mass, dtype=torch.float64))
        nI = torch.as_tensor(rm.norm_body_inertia, dtype=torch.float64)

        def grads(lo, hi):
            n = hi - lo
            sub = {k: (v[:, lo:hi] if k in ("torques", "res_f", "refs") else v[lo:hi]).contiguous() for k, v in d.items()}
            sub["target_ke"] = ke.detach()[None].repeat(n, 1)
            m = mass.detach()[None].repeat(n, 1)
            sub["body_inv_mass"], sub["body_inertia"] = 1.0 / m, nI[None] * m[..., None, None]
            sub["body_inv_inertia"] = torch.linalg.inv(sub["body_inertia"])
            cpu = CpuRollout(rm)
            pos, vel = cpu.forward(sub, 5e-4, stride, F)
            g = cpu.backward(pos, vel * 0.1)
            # chain rule of dp_model.py:725-730 onto the shared mass
            mm = m.clone().requires_grad_(True)
            I = nI[None] * mm[..., None, None]
            obj = ((1.0 / mm) * g["body_inv_mass"]).sum() + (I * g["body_inertia"]).sum() + \
                  (torch.linalg.inv(I) * g["body_inv_inertia"]).sum()
            gm, = torch.autograd.grad(obj, mm)
            return g["target_ke"].sum(0), gm.sum(0)

        lo, hi = shard_range(bs, rank, world)
        assert shard_envs(d["refs"].reshape(T, -1), bs, rank, world, env_dim=1).shape == (T, (hi - lo) * rm.nqd)
        ke.grad, mass.grad = grads(lo, hi)
        pg = PackedGrads([ke, mass])
        assert pg.numel() == rm.nqd + rm.nb
        pg.all_reduce()
        full_ke, full_m = grads(0, bs)
        ok = bool(torch.allclose(ke.grad, full_ke, rtol=1e-10, atol=1e-12) and
                  torch.allclose(mass.grad, full_m, rtol=1e-10, atol=1e-12))
        q.put((rank, ok, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_and_balance():
    from ppr_diffphys_b200.dist import shard_range
    for n in (1, 5, 64, 65537):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sharded_gradients_equal_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[2] for r in res) == [(0, 3), (3, 5)]
    assert all(r[1] for r in res)
