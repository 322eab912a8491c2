"""Loader of the C-ABI CUDA library (include/ppr_b200.h). There is NO CPU fallback: if libppr_b200.so is
missing or cannot be loaded every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

from ._capi import ModelDesc, RolloutIO

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PPR_B200_LIB") or os.path.join(_HERE, "libppr_b200.so")  # env override: A/B builds
_lib = None

_vp, _i64, _f32 = C.c_void_p, C.c_int64, C.c_float

EXPORTS = [
    "ppr_version", "ppr_model_create", "ppr_model_destroy", "ppr_model_set_joint_X_p", "ppr_model_set_joint_X_p_env", "ppr_model_set_attach",
    "ppr_model_set_gravity", "ppr_model_set_ground", "ppr_model_set_checkpoint_every", "ppr_model_set_latency_envs",
    "ppr_model_latency_envs", "ppr_model_set_team_envs", "ppr_model_team_envs", "ppr_model_envs_per_group", "ppr_model_group_threads", "ppr_fk_forward", "ppr_fk_backward",
    "ppr_rollout_workspace_bytes", "ppr_rollout_forward", "ppr_rollout_backward", "ppr_se3_loss_forward",
    "ppr_se3_loss_backward", "ppr_frame_compose_forward", "ppr_frame_compose_backward", "ppr_launch_count",
    "ppr_rollout_shared_grad_floats", "ppr_rollout_reduce_scratch_bytes", "ppr_rollout_backward_shared",
    "ppr_refs_from_frames", "ppr_refs_from_frames_backward", "ppr_rollout_forward_ex", "ppr_rollout_backward_ex",
]


class PprError(RuntimeError):
    pass


def _declare(lib):
    lib.ppr_version.restype = C.c_char_p
    lib.ppr_version.argtypes = []
    lib.ppr_model_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(_vp)]
    lib.ppr_model_destroy.argtypes = [_vp]
    lib.ppr_model_set_joint_X_p.argtypes = [_vp, _vp, _vp]
    lib.ppr_model_set_attach.argtypes = [_vp, _f32, _f32]
    lib.ppr_model_set_gravity.argtypes = [_vp, C.POINTER(C.c_float)]
    lib.ppr_model_set_ground.argtypes = [_vp, C.c_int32]
    lib.ppr_model_set_checkpoint_every.argtypes = [_vp, C.c_int32]
    lib.ppr_frame_compose_forward.argtypes = [_i64, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.ppr_frame_compose_backward.argtypes = [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.ppr_se3_loss_forward.argtypes = [_i64, C.c_int32, _vp, _vp, _f32, _vp, _vp]
    lib.ppr_se3_loss_backward.argtypes = [_i64, C.c_int32, _vp, _vp, _f32, _vp, _vp, _vp, _vp]
    lib.ppr_model_set_joint_X_p_env.argtypes = [_vp, _vp, _i64]
    lib.ppr_model_set_latency_envs.argtypes = [_vp, _i64]
    lib.ppr_model_latency_envs.argtypes = [_vp]
    lib.ppr_model_latency_envs.restype = _i64
    lib.ppr_model_set_team_envs.argtypes = [_vp, _i64]
    lib.ppr_model_team_envs.argtypes = [_vp]
    lib.ppr_model_team_envs.restype = _i64
    lib.ppr_model_envs_per_group.argtypes = [_vp]
    lib.ppr_model_group_threads.argtypes = [_vp]
    lib.ppr_fk_forward.argtypes = [_vp, _i64] + [_vp] * 5
    lib.ppr_fk_backward.argtypes = [_vp, _i64] + [_vp] * 7
    lib.ppr_rollout_workspace_bytes.restype = C.c_size_t
    lib.ppr_rollout_workspace_bytes.argtypes = [_vp, _i64, _i64]
    lib.ppr_rollout_forward.argtypes = [_vp, _i64, _i64, _i64, _f32, C.c_int32] + [_vp] * 14 + [_vp, C.c_size_t, _vp]
    lib.ppr_rollout_backward.argtypes = [_vp, _i64, _i64, _i64, _f32, C.c_int32] + [_vp] * 22 + [_vp, C.c_size_t, _vp]
    lib.ppr_rollout_shared_grad_floats.restype = _i64
    lib.ppr_rollout_shared_grad_floats.argtypes = [_vp]
    lib.ppr_rollout_reduce_scratch_bytes.restype = C.c_size_t
    lib.ppr_rollout_reduce_scratch_bytes.argtypes = [_vp, _i64]
    lib.ppr_rollout_backward_shared.argtypes = [_vp, _i64, _i64, _i64, _f32] + [_vp] * 18 + [_vp, C.c_size_t, _vp, C.c_size_t, _vp]
    lib.ppr_refs_from_frames.argtypes = [_i64, _i64, _i64, _i64, _vp, _vp, _vp]
    lib.ppr_refs_from_frames_backward.argtypes = [_i64, _i64, _i64, _i64, _vp, _vp, _vp]
    lib.ppr_rollout_forward_ex.argtypes = [_vp, C.POINTER(RolloutIO), _vp]
    lib.ppr_rollout_backward_ex.argtypes = [_vp, C.POINTER(RolloutIO), _vp]
    lib.ppr_launch_count.restype = C.c_int64
    lib.ppr_launch_count.argtypes = []
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:  # default
            fn.restype = C.c_int


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PprError("CUDA library %s not built -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


_ERR = {-1: "PPR_E_ARG (null / inconsistent argument)", -2: "PPR_E_SHAPE (unsupported size)",
        -3: "PPR_E_HANDLE (bad model handle)", -4: "PPR_E_WORKSPACE (workspace too small)"}


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise PprError("%s failed: %s" % (what, _ERR.get(rc, str(rc))))
    raise PprError("%s failed: cudaError_t %d" % (what, rc))


def launch_count():
    return int(lib().ppr_launch_count())
