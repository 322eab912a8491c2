"""ctypes front-end of the CPU port (oracle/cpu_port/ppr_cpu.cpp) -- test / baseline infrastructure only.

Same [bs,...] tensor layout as oracle.sim_oracle.rollout; float32 or float64 chosen by the input dtype."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import torch

from ppr_diffphys_b200._capi import make_desc

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libppr_cpu.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def num_threads():
    return int(lib().ppr_cpu_num_threads())


def use_all_cores():
    """Use every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().ppr_cpu_set_num_threads(int(n))
    return num_threads()


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _suf(dtype):
    return {torch.float32: "f32", torch.float64: "f64"}[dtype]


class CpuRollout:
    """Forward + hand-written reverse sweep on host cores."""

    def __init__(self, rm):
        self.rm = rm
        self.desc, self._keep = make_desc(rm)

    def forward(self, d, dt, stride, F, want_forces=False):
        rm = self.rm
        dtype = d["q_init"].dtype
        bs = d["q_init"].shape[0]
        T = stride * (F - 1) + 1
        c = {k: (v.contiguous() if v is not None else None) for k, v in d.items()}
        pos = torch.empty(F, bs, rm.nb, 7, dtype=dtype)
        vel = torch.empty(F, bs, rm.nb, 6, dtype=dtype)
        grf = torch.zeros(F, bs, rm.nb, 6, dtype=dtype) if want_forces else None
        jaf = torch.zeros(F, bs, rm.nb, 6, dtype=dtype) if want_forces else None
        states = torch.empty(T, bs, rm.nb, 13, dtype=dtype)
        fn = getattr(lib(), "ppr_cpu_rollout_forward_" + _suf(dtype))
        rc = fn(C.byref(self.desc), C.c_int64(bs), C.c_int64(T), C.c_int64(stride), C.c_double(dt), _p(c["q_init"]),
                _p(c["qd_init"]), _p(c.get("torques")), _p(c.get("res_f")), _p(c["refs"]), _p(c["target_ke"]),
                _p(c["target_kd"]), _p(c["body_inv_mass"]), _p(c["body_inertia"]), _p(c["body_inv_inertia"]),
                _p(pos), _p(vel), _p(grf), _p(jaf), _p(states))
        assert rc == 0
        self._saved = (c, dt, stride, F, states)
        return (pos, vel, grf, jaf) if want_forces else (pos, vel)

    def backward(self, adj_pos, adj_vel):
        rm = self.rm
        c, dt, stride, F, states = self._saved
        dtype = adj_pos.dtype
        bs = c["q_init"].shape[0]
        T = stride * (F - 1) + 1
        z = lambda *s: torch.zeros(*s, dtype=dtype)
        out = dict(q_init=z(bs, rm.nq), qd_init=z(bs, rm.nqd),
                   torques=z(T, bs, rm.nqd) if c.get("torques") is not None else None,
                   res_f=z(T, bs, rm.nb, 6) if c.get("res_f") is not None else None,
                   refs=z(T, bs, rm.nqd), target_ke=z(bs, rm.nqd), target_kd=z(bs, rm.nqd),
                   body_inv_mass=z(bs, rm.nb), body_inertia=z(bs, rm.nb, 3, 3), body_inv_inertia=z(bs, rm.nb, 3, 3))
        fn = getattr(lib(), "ppr_cpu_rollout_backward_" + _suf(dtype))
        rc = fn(C.byref(self.desc), C.c_int64(bs), C.c_int64(T), C.c_int64(stride), C.c_double(dt), _p(c["q_init"]),
                _p(c["qd_init"]), _p(c.get("torques")), _p(c.get("res_f")), _p(c["refs"]), _p(c["target_ke"]),
                _p(c["target_kd"]), _p(c["body_inv_mass"]), _p(c["body_inertia"]), _p(c["body_inv_inertia"]),
                _p(states), _p(adj_pos.contiguous()), _p(adj_vel.contiguous()), _p(out["q_init"]), _p(out["qd_init"]),
                _p(out["torques"]), _p(out["res_f"]), _p(out["refs"]), _p(out["target_ke"]), _p(out["target_kd"]),
                _p(out["body_inv_mass"]), _p(out["body_inertia"]), _p(out["body_inv_inertia"]))
        assert rc == 0
        return out

    def fk(self, jq, jqd):
        rm = self.rm
        n = jq.shape[0]
        bq = torch.empty(n, rm.nb, 7, dtype=jq.dtype)
        bqd = torch.empty(n, rm.nb, 6, dtype=jq.dtype)
        fn = getattr(lib(), "ppr_cpu_fk_forward_" + _suf(jq.dtype))
        assert fn(C.byref(self.desc), C.c_int64(n), _p(jq.contiguous()), _p(jqd.contiguous()), _p(bq), _p(bqd)) == 0
        return bq, bqd

    def fk_backward(self, jq, jqd, adj_q, adj_qd):
        n = jq.shape[0]
        ajq, ajqd = torch.zeros_like(jq), torch.zeros_like(jqd)
        fn = getattr(lib(), "ppr_cpu_fk_backward_" + _suf(jq.dtype))
        assert fn(C.byref(self.desc), C.c_int64(n), _p(jq.contiguous()), _p(jqd.contiguous()), _p(adj_q.contiguous()),
                  _p(adj_qd.contiguous()), _p(ajq), _p(ajqd)) == 0
        return ajq, ajqd
