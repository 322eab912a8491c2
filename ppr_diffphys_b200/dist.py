"""Multi-GPU plumbing: environments (robot x time-window instances) are independent in forward and backward
(the reference already builds them as disjoint articulations, dp_model.py:384-386), so they shard across ranks
with NO data-path collective.  The only exchange is ONE all-reduce per optimisation step of the packed
shared-parameter gradients (target_ke[nqd], target_kd[nqd], body_mass[nb], optionally global_q[7]) -- <= ~200
floats, latency-bound on NVLink.  The reference has no distributed code at all (main.py:15-16 flags are unused).

Works with any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(num_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous env range [lo, hi) of `rank`; sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(num_envs, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_envs(t: torch.Tensor, num_envs: int, rank: int, world: int, env_dim: int = 0) -> torch.Tensor:
    """Slice the env dimension of a tensor laid out [..., num_envs * k, ...] along `env_dim` (flattened per-env
    layout of the reference, dp_model.py:563-572)."""
    lo, hi = shard_range(num_envs, rank, world)
    per = t.shape[env_dim] // num_envs
    assert per * num_envs == t.shape[env_dim], "env dimension is not a multiple of num_envs"
    return t.narrow(env_dim, lo * per, (hi - lo) * per)


class PackedGrads:
    """Packs the .grad of a few small shared parameters into one flat buffer, all-reduces it ONCE (sum) and
    scatters the result back -- the single collective of an optimisation step."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        self.buf = torch.zeros(sum(self.sizes), device=self.params[0].device, dtype=self.params[0].dtype)

    def numel(self):
        return int(self.buf.numel())

    def pack(self):
        o = 0
        for p, n in zip(self.params, self.sizes):
            self.buf[o:o + n] = 0 if p.grad is None else p.grad.reshape(-1)
            o += n
        return self.buf

    def unpack(self):
        o = 0
        for p, n in zip(self.params, self.sizes):
            g = self.buf[o:o + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            o += n

    def all_reduce(self, average_over: int = 0):
        """sum over ranks (divide by `average_over` if > 0, e.g. the global env count for a mean loss)."""
        self.pack()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
        if average_over:
            self.buf /= float(average_over)
        self.unpack()
        return self.buf
