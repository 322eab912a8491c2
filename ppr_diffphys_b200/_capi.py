"""ctypes mirror of include/ppr_b200.h (struct ppr_model_desc) -- no torch, no CUDA needed to import."""
from __future__ import annotations

import ctypes as C

import numpy as np

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


class ModelDesc(C.Structure):
    _fields_ = [
        ("nb", C.c_int32), ("nq", C.c_int32), ("nqd", C.c_int32), ("nc", C.c_int32), ("nshape", C.c_int32),
        ("joint_type", _i32p), ("joint_parent", _i32p), ("joint_q_start", _i32p), ("joint_qd_start", _i32p),
        ("joint_X_p", _f32p), ("joint_X_c", _f32p), ("joint_axis", _f32p),
        ("joint_limit_lower", _f32p), ("joint_limit_upper", _f32p), ("joint_limit_ke", _f32p),
        ("joint_limit_kd", _f32p), ("body_com", _f32p),
        ("contact_body", _i32p), ("contact_point", _f32p), ("contact_dist", _f32p), ("contact_material", _i32p),
        ("shape_materials", _f32p),
        ("gravity", C.c_float * 3), ("joint_attach_ke", C.c_float), ("joint_attach_kd", C.c_float),
    ]


def make_desc(rm):
    """RobotModel -> (ModelDesc, keepalive list of the numpy arrays the pointers refer to)."""
    keep = []

    def f32(a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
        keep.append(a)
        return a.ctypes.data_as(_f32p)

    def i32(a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.int32))
        keep.append(a)
        return a.ctypes.data_as(_i32p)

    d = ModelDesc()
    d.nb, d.nq, d.nqd, d.nc, d.nshape = rm.nb, rm.nq, rm.nqd, rm.nc, int(rm.shape_materials.shape[0])
    d.joint_type, d.joint_parent = i32(rm.joint_type), i32(rm.joint_parent)
    d.joint_q_start, d.joint_qd_start = i32(rm.joint_q_start), i32(rm.joint_qd_start)
    d.joint_X_p, d.joint_X_c, d.joint_axis = f32(rm.joint_X_p), f32(rm.joint_X_c), f32(rm.joint_axis)
    d.joint_limit_lower, d.joint_limit_upper = f32(rm.joint_limit_lower), f32(rm.joint_limit_upper)
    d.joint_limit_ke, d.joint_limit_kd = f32(rm.joint_limit_ke), f32(rm.joint_limit_kd)
    d.body_com = f32(rm.body_com)
    d.contact_body, d.contact_point = i32(rm.contact_body), f32(rm.contact_point)
    d.contact_dist, d.contact_material = f32(rm.contact_dist), i32(rm.contact_material)
    d.shape_materials = f32(rm.shape_materials)
    g = np.asarray(rm.gravity, dtype=np.float32)
    d.gravity[0], d.gravity[1], d.gravity[2] = float(g[0]), float(g[1]), float(g[2])
    d.joint_attach_ke, d.joint_attach_kd = float(rm.joint_attach_ke), float(rm.joint_attach_kd)
    return d, keep


_vp = C.c_void_p


class RolloutIO(C.Structure):
    """struct ppr_rollout_io (include/ppr_b200.h): every option of the rollout; pointers are raw device addresses."""
    _fields_ = [
        ("bs", C.c_int64), ("nsteps", C.c_int64), ("frame_stride", C.c_int64), ("dt", C.c_float),
        ("shared_params", C.c_int32),
        ("q_init", _vp), ("qd_init", _vp), ("torques", _vp), ("res_f", _vp), ("refs", _vp), ("target_ke", _vp),
        ("target_kd", _vp), ("body_inv_mass", _vp), ("body_inertia", _vp), ("body_inv_inertia", _vp),
        ("out_pos", _vp), ("out_vel", _vp), ("out_grf", _vp), ("out_jaf", _vp),
        ("workspace", _vp), ("workspace_bytes", C.c_size_t),
        ("target_pos", _vp), ("rot_ratio", C.c_float), ("loss_pos", _vp), ("adj_loss_pos", _vp),
        ("adj_target_pos", _vp),
        ("adj_out_pos", _vp), ("adj_out_vel", _vp),
        ("adj_q_init", _vp), ("adj_qd_init", _vp), ("adj_torques", _vp), ("adj_res_f", _vp), ("adj_refs", _vp),
        ("adj_target_ke", _vp), ("adj_target_kd", _vp), ("adj_body_inv_mass", _vp), ("adj_body_inertia", _vp),
        ("adj_body_inv_inertia", _vp),
        ("adj_shared", _vp), ("reduce_scratch", _vp), ("reduce_scratch_bytes", C.c_size_t),
    ]
