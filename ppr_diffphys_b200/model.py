"""Model compiler: URDF -> static per-articulation arrays the rollout kernels read.

Restates, for ONE articulation, what the reference obtains from
``parse_urdf`` + Warp's ``ModelBuilder`` + ``Model.collide``:

* /root/reference/diffphys/import_urdf.py:106-291  (link/joint rules, ``_R/_P/_Y`` collapse
  into JOINT_COMPOUND, density>0 => URDF inertials ignored)
* /root/reference/diffphys/dp_model.py:76-205       (per-robot constants, feet scaling, inertia
  normalisation, PD gain vectors)
* /root/reference/diffphys/dp_model.py:384-401      (``env.ground = True``, ``collide()`` once)

The Warp 0.7.2 builder itself is a third-party package that is not vendored under
/root/reference; its rules (shape mass, parallel-axis merge, contact generation) are
restated from its published behaviour and are flagged UNVERIFIED in DESIGN.md.  Every
quantity it produces is an explicit input of the kernels, so parity tests feed identical
arrays to the oracle and to the CUDA path; ``RobotModel.load`` also accepts a ``.npz``
dumped from a real Warp ``Model``.

The reference replicates these arrays ``num_envs`` times in device memory
(dp_model.py:384-386); here ONE copy is uploaded and shared by every environment.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field, asdict
from typing import Dict, List, Optional

import numpy as np

from . import urdf as _urdf

# Warp joint-type enum values (warp.sim.model, 0.7.x)
JOINT_PRISMATIC, JOINT_REVOLUTE, JOINT_BALL, JOINT_FIXED, JOINT_FREE, JOINT_COMPOUND, JOINT_UNIVERSAL = range(7)
_DOF = {JOINT_PRISMATIC: (1, 1), JOINT_REVOLUTE: (1, 1), JOINT_BALL: (3, 4), JOINT_FIXED: (0, 0),
        JOINT_FREE: (6, 7), JOINT_COMPOUND: (3, 3), JOINT_UNIVERSAL: (2, 2)}

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


# ----------------------------------------------------------------------------- small math
def quat_rpy(roll, pitch, yaw):
    cy, sy = math.cos(yaw * 0.5), math.sin(yaw * 0.5)
    cr, sr = math.cos(roll * 0.5), math.sin(roll * 0.5)
    cp, sp = math.cos(pitch * 0.5), math.sin(pitch * 0.5)
    w = cy * cr * cp + sy * sr * sp
    x = cy * sr * cp - sy * cr * sp
    y = cy * cr * sp + sy * sr * cp
    z = sy * cr * cp - cy * sr * sp
    return np.array([x, y, z, w])


def quat_from_axis_angle(axis, angle):
    a = np.asarray(axis, dtype=np.float64)
    return np.array([*(a * math.sin(angle * 0.5)), math.cos(angle * 0.5)])


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + bw * ax + ay * bz - az * by,
                     aw * by + bw * ay + az * bx - ax * bz,
                     aw * bz + bw * az + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def quat_to_matrix(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _steiner(m, I, p, q):
    R = quat_to_matrix(q)
    return R @ I @ R.T + m * (np.dot(p, p) * np.eye(3) - np.outer(p, p))


def mesh_mass_inertia(V: np.ndarray, F: np.ndarray):
    """Unit-density mass and inertia of a closed triangle mesh about its vertex mean, by signed
    tetrahedra with the 4-point order-2 quadrature (weight 1/4, alpha = sqrt(5)/5)."""
    com = V.mean(0)
    p, q, r = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    vol = np.einsum("ij,ij->i", p - com, np.cross(q - com, r - com)) / 6.0
    mid = (com[None] + p + q + r) / 4.0
    alpha = math.sqrt(5.0) / 5.0
    I = np.zeros((3, 3))
    for vert in (p, q, r, np.broadcast_to(com, p.shape)):
        d = mid + (vert - mid) * alpha - com
        dd = np.einsum("ij,ij->i", d, d)
        I += 0.25 * (np.einsum("i,i->", vol, dd) * np.eye(3) - np.einsum("i,ij,ik->jk", vol, d, d))
    return float(vol.sum()), I, com


# ----------------------------------------------------------------------------- builder
class ArticulationBuilder:
    """One-articulation subset of Warp's ModelBuilder (add_body / add_shape_* / collide)."""

    def __init__(self):
        self.body_mass: List[float] = []
        self.body_inertia: List[np.ndarray] = []
        self.body_com: List[np.ndarray] = []
        self.body_name: List[str] = []
        self.joint_type: List[int] = []
        self.joint_parent: List[int] = []
        self.joint_X_p: List[np.ndarray] = []
        self.joint_X_c: List[np.ndarray] = []
        self.joint_axis: List[np.ndarray] = []
        self.joint_q_start: List[int] = []
        self.joint_qd_start: List[int] = []
        self.joint_q: List[float] = []
        self.joint_limit_lower: List[float] = []
        self.joint_limit_upper: List[float] = []
        self.joint_limit_ke: List[float] = []
        self.joint_limit_kd: List[float] = []
        self.joint_target_ke: List[float] = []
        self.joint_target_kd: List[float] = []
        self.nqd = 0
        self.shape_body: List[int] = []
        self.shape_pos: List[np.ndarray] = []
        self.shape_rot: List[np.ndarray] = []
        self.shape_geo_type: List[str] = []
        self.shape_geo_scale: List[tuple] = []
        self.shape_geo_src: List[Optional[np.ndarray]] = []
        self.shape_materials: List[tuple] = []

    def add_body(self, parent, joint_type, joint_xform=None, joint_xform_child=None, joint_axis=(0.0, 0.0, 0.0),
                 lower=-1e3, upper=1e3, limit_ke=0.0, limit_kd=0.0, target_ke=0.0, target_kd=0.0,
                 armature=0.0, com=None, I_m=None, m=0.0, name=""):
        ident = np.array([0, 0, 0, 0, 0, 0, 1.0])
        dof, coord = _DOF[joint_type]
        self.body_mass.append(float(m))
        self.body_inertia.append((np.zeros((3, 3)) if I_m is None else np.asarray(I_m, float)) + np.eye(3) * armature)
        self.body_com.append(np.zeros(3) if com is None else np.asarray(com, float))
        self.body_name.append(name)
        self.joint_type.append(joint_type)
        self.joint_parent.append(parent)
        self.joint_X_p.append(ident.copy() if joint_xform is None else np.asarray(joint_xform, float))
        self.joint_X_c.append(ident.copy() if joint_xform_child is None else np.asarray(joint_xform_child, float))
        self.joint_axis.append(np.asarray(joint_axis, float))
        self.joint_q_start.append(len(self.joint_q))
        self.joint_qd_start.append(self.nqd)
        self.joint_q.extend([0.0] * coord)
        if joint_type == JOINT_FREE:
            self.joint_q[-1] = 1.0
        for _ in range(dof):
            self.joint_limit_lower.append(lower)
            self.joint_limit_upper.append(upper)
            self.joint_limit_ke.append(limit_ke)
            self.joint_limit_kd.append(limit_kd)
            self.joint_target_ke.append(target_ke)
            self.joint_target_kd.append(target_kd)
        self.nqd += dof
        return len(self.body_mass) - 1

    def _update_body_mass(self, i, m, I, p, q):
        new_mass = self.body_mass[i] + m
        if new_mass == 0.0:
            return
        new_com = (self.body_com[i] * self.body_mass[i] + p * m) / new_mass
        ident = np.array([0, 0, 0, 1.0])
        self.body_inertia[i] = (_steiner(self.body_mass[i], self.body_inertia[i], new_com - self.body_com[i], ident)
                                + _steiner(m, I, new_com - p, q))
        self.body_mass[i] = new_mass
        self.body_com[i] = new_com

    def _add_shape(self, body, pos, rot, kind, scale, src, m, I, mat):
        self.shape_body.append(body)
        self.shape_pos.append(np.asarray(pos, float))
        self.shape_rot.append(np.asarray(rot, float))
        self.shape_geo_type.append(kind)
        self.shape_geo_scale.append(tuple(scale))
        self.shape_geo_src.append(src)
        self.shape_materials.append(tuple(mat))
        self._update_body_mass(body, m, I, np.asarray(pos, float), np.asarray(rot, float))

    def add_shape_box(self, body, pos, rot, hx, hy, hz, density, mat):
        m = density * 8.0 * hx * hy * hz
        I = m / 12.0 * np.diag([(2 * hy) ** 2 + (2 * hz) ** 2, (2 * hx) ** 2 + (2 * hz) ** 2, (2 * hx) ** 2 + (2 * hy) ** 2])
        self._add_shape(body, pos, rot, "box", (hx, hy, hz), None, m, I, mat)

    def add_shape_sphere(self, body, pos, rot, radius, density, mat):
        m = density * 4.0 / 3.0 * math.pi * radius ** 3
        I = 2.0 / 5.0 * m * radius * radius * np.eye(3)
        self._add_shape(body, pos, rot, "sphere", (radius, 0.0, 0.0), None, m, I, mat)

    def add_shape_capsule(self, body, pos, rot, radius, half_width, density, mat):
        ms = density * 4.0 / 3.0 * math.pi * radius ** 3
        mc = density * math.pi * radius * radius * 2.0 * half_width
        m = ms + mc
        Ia = mc * (0.5 * radius ** 2) + ms * (0.4 * radius ** 2)
        Ib = (mc * (0.25 * radius ** 2 + (1.0 / 3.0) * half_width ** 2)
              + ms * (0.4 * radius ** 2 + 0.75 * half_width * radius + half_width ** 2))
        self._add_shape(body, pos, rot, "capsule", (radius, half_width, 0.0), None, m, np.diag([Ia, Ib, Ib]), mat)

    def add_shape_mesh(self, body, pos, rot, V, F, density, mat, scale=1.0):
        vol, I, _ = mesh_mass_inertia(V, F)
        self._add_shape(body, pos, rot, "mesh", (scale, scale, scale), V, density * vol * scale ** 3,
                        density * I * scale ** 5, mat)

    def collide(self):
        """Static ground-contact candidates in body frame: sphere 1 pt (dist=r), capsule 2 pts on +-x
        (dist=r), box 8 corners, mesh every vertex (dist=0)."""
        cb, cp, cd, cm = [], [], [], []
        for s, body in enumerate(self.shape_body):
            R = quat_to_matrix(self.shape_rot[s])
            pos = self.shape_pos[s]
            kind, sc = self.shape_geo_type[s], self.shape_geo_scale[s]
            if kind == "sphere":
                pts, dist = np.zeros((1, 3)), sc[0]
            elif kind == "capsule":
                pts, dist = np.array([[-sc[1], 0, 0], [sc[1], 0, 0.0]]), sc[0]
            elif kind == "box":
                pts = np.array([[sx * sc[0], sy * sc[1], sz * sc[2]]
                                for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=float)
                dist = 0.0
            else:
                pts, dist = self.shape_geo_src[s] * sc[0], 0.0
            for p in pts:
                cb.append(body)
                cp.append(pos + R @ p)
                cd.append(dist)
                cm.append(s)
        return (np.asarray(cb, np.int32), np.asarray(cp, np.float64).reshape(-1, 3),
                np.asarray(cd, np.float64), np.asarray(cm, np.int32))


def _add_collisions(b: ArticulationBuilder, body, collisions, density, mat):
    for c in collisions:
        rot = quat_rpy(*c.rpy)
        if c.kind == "box":
            b.add_shape_box(body, c.xyz, rot, c.size[0] * 0.5, c.size[1] * 0.5, c.size[2] * 0.5, density, mat)
        elif c.kind == "sphere":
            b.add_shape_sphere(body, c.xyz, rot, c.radius, density, mat)
        elif c.kind == "cylinder":
            r = quat_from_axis_angle((0.0, 1.0, 0.0), math.pi * 0.5)
            b.add_shape_capsule(body, c.xyz, quat_mul(rot, r), c.radius, c.length * 0.5, density, mat)
        elif c.kind == "mesh":
            for V, F in c.meshes:
                b.add_shape_mesh(body, c.xyz, rot, V, F, density, mat)


def parse_urdf(filename, builder: ArticulationBuilder, density=1000.0, armature=0.01, stiffness=220.0,
               damping=2.0, shape_ke=1e4, shape_kd=0.0, shape_kf=1e2, shape_mu=1.0, limit_ke=0.0, limit_kd=0.0):
    """Floating-base import with the reference's rules (import_urdf.py:106-291, ``floating=True``,
    ``density>0`` branch)."""
    robot = _urdf.load_urdf(filename)
    mat = (shape_ke, shape_kd, shape_kf, shape_mu)
    link_index: Dict[str, int] = {}
    root = builder.add_body(-1, JOINT_FREE, armature=armature, name=robot.links[0].name)
    _add_collisions(builder, root, robot.links[0].collisions, density, mat)
    link_index[robot.links[0].name] = root
    link_map = robot.link_map
    for j in robot.joints:
        jt, axis = None, np.zeros(3)
        if j.joint_type in ("revolute", "continuous"):
            jt, axis = JOINT_REVOLUTE, j.axis
        elif j.joint_type == "prismatic":
            jt, axis = JOINT_PRISMATIC, j.axis
        elif j.joint_type == "fixed":
            jt = JOINT_FIXED
        elif j.joint_type == "floating":
            jt = JOINT_FREE
        child = j.child
        if j.name[-2:] == "_R":
            jt = JOINT_COMPOUND
            child = child[:-2] + "_Y"
        elif j.name[-2:] in ("_P", "_Y"):
            continue
        parent = link_index.get(j.parent, root)
        lower = j.lower if j.lower is not None else -1e3
        upper = j.upper if j.upper is not None else 1e3
        if j.damping:
            damping = j.damping
        xf = np.concatenate([j.xyz, quat_rpy(*j.rpy)])
        if jt == JOINT_COMPOUND:
            link = builder.add_body(parent, jt, joint_xform=xf, joint_xform_child=np.array([0, 0, 0, 0, 0, 0, 1.0]),
                                    lower=lower, upper=upper, limit_ke=limit_ke, limit_kd=limit_kd,
                                    target_ke=stiffness, target_kd=damping, armature=armature, name=child)
        else:
            link = builder.add_body(parent, jt, joint_xform=xf, joint_axis=axis, lower=lower, upper=upper,
                                    limit_ke=limit_ke, limit_kd=limit_kd, target_ke=stiffness, target_kd=damping,
                                    armature=armature, name=child)
        _add_collisions(builder, link, link_map[child].collisions, density, mat)
        link_index[child] = link
    return robot


# ----------------------------------------------------------------------------- compiled model
_ARRAY_FIELDS = ("joint_type", "joint_parent", "joint_X_p", "joint_X_c", "joint_axis", "joint_q_start",
                 "joint_qd_start", "joint_limit_lower", "joint_limit_upper", "joint_limit_ke", "joint_limit_kd",
                 "joint_target_ke", "joint_target_kd", "body_com", "body_mass", "norm_body_inertia",
                 "contact_body", "contact_point", "contact_dist", "contact_material", "shape_materials",
                 "gravity", "joint_q_rest")


@dataclass
class RobotModel:
    """Static arrays of one articulation (names follow the Warp ``Model`` attributes the
    reference kernels read: integrator_euler.py:497-504,519-538,603-611)."""
    name: str
    joint_type: np.ndarray          # [nb] int32
    joint_parent: np.ndarray        # [nb] int32
    joint_X_p: np.ndarray           # [nb,7] f32 (p, q xyzw)
    joint_X_c: np.ndarray           # [nb,7] f32
    joint_axis: np.ndarray          # [nb,3] f32
    joint_q_start: np.ndarray       # [nb] int32
    joint_qd_start: np.ndarray      # [nb] int32
    joint_limit_lower: np.ndarray   # [nqd]
    joint_limit_upper: np.ndarray   # [nqd]
    joint_limit_ke: np.ndarray      # [nqd]
    joint_limit_kd: np.ndarray      # [nqd]
    joint_target_ke: np.ndarray     # [nqd] default PD gains (0 on the 6 root dofs)
    joint_target_kd: np.ndarray     # [nqd]
    body_com: np.ndarray            # [nb,3]
    body_mass: np.ndarray           # [nb] initial value of the learnable mass
    norm_body_inertia: np.ndarray   # [nb,3,3] inertia / mass (dp_model.py:179-196)
    contact_body: np.ndarray        # [nc] int32
    contact_point: np.ndarray       # [nc,3]
    contact_dist: np.ndarray        # [nc]
    contact_material: np.ndarray    # [nc] int32 -> row of shape_materials
    shape_materials: np.ndarray     # [nshape,4] (ke, kd, kf, mu)
    gravity: np.ndarray             # [3]
    joint_q_rest: np.ndarray        # [nq] builder joint_q (root at spawn xform)
    joint_attach_ke: float = 0.0
    joint_attach_kd: float = 0.0
    body_names: List[str] = field(default_factory=list)

    def as_builder_view(self):
        """The four lists ``phys_model.__init__`` reads from its Warp ``articulation_builder`` to initialise the
        ``nn.Parameter``s (dp_model.py:198-222) -- here already post-processed (inertia / mass, PD vectors)."""
        from types import SimpleNamespace
        return SimpleNamespace(joint_target_ke=[float(x) for x in self.joint_target_ke],
                               joint_target_kd=[float(x) for x in self.joint_target_kd],
                               body_mass=[float(x) for x in self.body_mass],
                               body_inertia=np.asarray(self.norm_body_inertia, dtype=np.float32).copy())

    @property
    def nb(self):
        return int(self.joint_type.shape[0])

    @property
    def nq(self):
        return int(self.joint_q_rest.shape[0])

    @property
    def nqd(self):
        return int(self.joint_target_ke.shape[0])

    @property
    def nc(self):
        return int(self.contact_body.shape[0])

    def save(self, path):
        d = {k: getattr(self, k) for k in _ARRAY_FIELDS}
        np.savez_compressed(path, name=self.name, joint_attach_ke=self.joint_attach_ke,
                            joint_attach_kd=self.joint_attach_kd, body_names=np.array(self.body_names), **d)

    @staticmethod
    def load(path) -> "RobotModel":
        z = np.load(path, allow_pickle=False)
        kw = {k: z[k] for k in _ARRAY_FIELDS}
        return RobotModel(name=str(z["name"]), joint_attach_ke=float(z["joint_attach_ke"]),
                          joint_attach_kd=float(z["joint_attach_kd"]),
                          body_names=[str(s) for s in z["body_names"]], **kw)


# per-robot constants, dp_model.py:83-119
ROBOT_PRESETS = {
    "laikago": dict(urdf="laikago/laikago.urdf", attach_ke=16000.0, attach_kd=200.0, kp=220.0, kd=2.0,
                    shape_ke=1e4, shape_kd=0.0, kp_links=None),
    "quad": dict(urdf="quad.urdf", attach_ke=8000.0, attach_kd=200.0, kp=660.0, kd=5.0, shape_ke=1e4, shape_kd=0.0,
                 kp_links=["link_155_Vorderpfote_R_Y", "link_150_Vorderpfote_L_Y", "link_170_Pfote2_R_Y",
                           "link_165_Pfote2_L_Y"]),
    "human": dict(urdf="human.urdf", attach_ke=8000.0, attach_kd=200.0, kp=660.0, kd=5.0, shape_ke=1e4, shape_kd=0.0,
                  kp_links=["link_24_mixamorig:RightFoot_Y", "link_19_mixamorig:LeftFoot_Y"]),
}


def compile_robot(name: str, urdf_root: str, gravity=(0.0, -9.80665, 0.0)) -> RobotModel:
    """URDF -> RobotModel with the reference's post-processing (dp_model.py:128-205)."""
    ps = ROBOT_PRESETS[name]
    b = ArticulationBuilder()
    parse_urdf(os.path.join(urdf_root, ps["urdf"]), b, density=1000.0, armature=0.01, stiffness=220.0, damping=2.0,
               shape_ke=ps["shape_ke"], shape_kd=ps["shape_kd"], shape_kf=1e2, shape_mu=1.0, limit_ke=0.0, limit_kd=0.0)
    # root spawn transform (dp_model.py:131-134)
    b.joint_q[0:7] = [0.0, 0.417, 0.0, 0.0, 0.0, 0.0, 1.0]
    nb = len(b.body_mass)
    if ps["kp_links"] is not None:
        # one box shape per body for these robots => shape index == body index (dp_model.py:151-191)
        assert len(b.shape_body) == nb and list(b.shape_body) == list(range(nb))
        for idx, nm in enumerate(b.body_name):
            if nm in ps["kp_links"]:
                b.shape_geo_scale[idx] = tuple(2.0 * s for s in b.shape_geo_scale[idx])
                b.body_mass[idx] *= 2 ** 3
                b.body_inertia[idx] = b.body_inertia[idx] * 2 ** 5
            b.body_inertia[idx] = b.body_inertia[idx] / b.body_mass[idx]
            w = 1e3 * float(np.prod(b.shape_geo_scale[idx]))
            b.body_mass[idx] = min(5.0, max(1.0, w))
    else:
        for idx in range(nb):
            b.body_inertia[idx] = b.body_inertia[idx] / b.body_mass[idx]
    nqd = b.nqd
    ke = [0.0] * 6 + [ps["kp"]] * (nqd - 6)
    kd = [0.0] * 6 + [ps["kd"]] * (nqd - 6)
    cb, cp, cd, cm = b.collide()
    f32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    i32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    return RobotModel(
        name=name, joint_type=i32(b.joint_type), joint_parent=i32(b.joint_parent), joint_X_p=f32(b.joint_X_p),
        joint_X_c=f32(b.joint_X_c), joint_axis=f32(b.joint_axis), joint_q_start=i32(b.joint_q_start),
        joint_qd_start=i32(b.joint_qd_start), joint_limit_lower=f32(b.joint_limit_lower),
        joint_limit_upper=f32(b.joint_limit_upper), joint_limit_ke=f32(b.joint_limit_ke),
        joint_limit_kd=f32(b.joint_limit_kd), joint_target_ke=f32(ke), joint_target_kd=f32(kd),
        body_com=f32(b.body_com), body_mass=f32(b.body_mass), norm_body_inertia=f32(b.body_inertia),
        contact_body=i32(cb), contact_point=f32(cp), contact_dist=f32(cd), contact_material=i32(cm),
        shape_materials=f32(b.shape_materials), gravity=f32(gravity), joint_q_rest=f32(b.joint_q),
        joint_attach_ke=ps["attach_ke"], joint_attach_kd=ps["attach_kd"], body_names=list(b.body_name))


def load_robot(name: str) -> RobotModel:
    """Load a pre-compiled robot from the package assets (made by tools/compile_assets.py)."""
    path = os.path.join(ASSET_DIR, "%s.npz" % name)
    if not os.path.exists(path):
        raise FileNotFoundError("compiled robot asset missing: %s (run tools/compile_assets.py)" % path)
    return RobotModel.load(path)
