"""Minimal URDF + collision-mesh reader (xml.etree only; no urdfpy / trimesh).

This is the setup-time producer of the static arrays the rollout kernels read.
It restates only what the reference's model import needs
(/root/reference/diffphys/import_urdf.py:23-103 consumes ``collision.origin``,
``geometry.{box,sphere,cylinder,mesh}``; :177-291 consumes joint name / type /
parent / child / origin / axis / limit).  urdfpy 0.0.22 + trimesh 3.9.43 are not
installable here, so mesh handling restates their documented behaviour:
an OBJ with several ``o``/``g`` groups is a scene of several meshes, every mesh
has coincident vertices merged.
"""
from __future__ import annotations

import os
import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np


@dataclass
class Collision:
    xyz: np.ndarray
    rpy: np.ndarray
    kind: str  # box | sphere | cylinder | mesh
    size: Optional[np.ndarray] = None  # box full extents
    radius: float = 0.0
    length: float = 0.0
    meshes: List[Tuple[np.ndarray, np.ndarray]] = field(default_factory=list)  # (V[n,3], F[m,3])


@dataclass
class Link:
    name: str
    collisions: List[Collision]
    inertial_xyz: np.ndarray
    inertial_mass: float
    inertial_I: np.ndarray


@dataclass
class Joint:
    name: str
    joint_type: str
    parent: str
    child: str
    xyz: np.ndarray
    rpy: np.ndarray
    axis: np.ndarray
    lower: Optional[float]
    upper: Optional[float]
    damping: Optional[float]


@dataclass
class Robot:
    name: str
    links: List[Link]
    joints: List[Joint]

    @property
    def link_map(self):
        return {l.name: l for l in self.links}


def _floats(s: Optional[str], n: int, default: float = 0.0) -> np.ndarray:
    if s is None:
        return np.full(n, default, dtype=np.float64)
    v = np.array([float(t) for t in s.split()], dtype=np.float64)
    assert v.shape[0] == n, (s, n)
    return v


def _merge_vertices(V: np.ndarray, F: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Merge coincident vertices (trimesh ``process=True`` behaviour), keep first-seen order,
    drop vertices no face references."""
    if len(V) == 0:
        return V, F
    used = np.zeros(len(V), dtype=bool)
    used[F.reshape(-1)] = True
    key = np.round(V * 1e8).astype(np.int64)
    _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    # stable order: order unique groups by first occurrence among *used* vertices
    group_used = np.zeros(len(first), dtype=bool)
    group_used[inv[used]] = True
    order = np.argsort(first, kind="stable")
    order = order[group_used[order]]
    remap = -np.ones(len(first), dtype=np.int64)
    remap[order] = np.arange(len(order))
    Vn = V[first[order]]
    Fn = remap[inv[F]]
    return Vn, Fn


def load_obj(path: str) -> List[Tuple[np.ndarray, np.ndarray]]:
    """Wavefront OBJ -> list of (V,F), one per ``o``/``g`` group that owns faces."""
    verts: List[List[float]] = []
    groups: List[List[List[int]]] = [[]]
    with open(path, "r", errors="ignore") as fh:
        for line in fh:
            if not line or line[0] == "#":
                continue
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif tok[0] in ("o", "g"):
                if groups[-1]:
                    groups.append([])
            elif tok[0] == "f":
                idx = []
                for t in tok[1:]:
                    i = int(t.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for k in range(1, len(idx) - 1):  # fan-triangulate polygons
                    groups[-1].append([idx[0], idx[k], idx[k + 1]])
    V = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    out = []
    for g in groups:
        if not g:
            continue
        F = np.asarray(g, dtype=np.int64)
        out.append(_merge_vertices(V, F))
    return out


def load_stl(path: str) -> List[Tuple[np.ndarray, np.ndarray]]:
    with open(path, "rb") as fh:
        data = fh.read()
    ntri = struct.unpack_from("<I", data, 80)[0] if len(data) >= 84 else 0
    if len(data) == 84 + 50 * ntri:  # binary
        rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                            count=ntri, offset=84)
        V = rec["v"].reshape(-1, 3).astype(np.float64)
    else:  # ascii
        pts = []
        for line in data.decode("ascii", errors="ignore").splitlines():
            tok = line.split()
            if len(tok) == 4 and tok[0] == "vertex":
                pts.append([float(tok[1]), float(tok[2]), float(tok[3])])
        V = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
    F = np.arange(len(V), dtype=np.int64).reshape(-1, 3)
    return [_merge_vertices(V, F)]


def load_mesh(path: str) -> List[Tuple[np.ndarray, np.ndarray]]:
    ext = os.path.splitext(path)[1].lower()
    if ext == ".obj":
        return load_obj(path)
    if ext == ".stl":
        return load_stl(path)
    raise ValueError("unsupported mesh format: %s" % path)


def _parse_collision(node: ET.Element, base_dir: str) -> Optional[Collision]:
    origin = node.find("origin")
    xyz = _floats(origin.get("xyz") if origin is not None else None, 3)
    rpy = _floats(origin.get("rpy") if origin is not None else None, 3)
    geo = node.find("geometry")
    if geo is None:
        return None
    box, sph, cyl, mesh = geo.find("box"), geo.find("sphere"), geo.find("cylinder"), geo.find("mesh")
    if box is not None:
        return Collision(xyz, rpy, "box", size=_floats(box.get("size"), 3))
    if sph is not None:
        return Collision(xyz, rpy, "sphere", radius=float(sph.get("radius")))
    if cyl is not None:
        return Collision(xyz, rpy, "cylinder", radius=float(cyl.get("radius")), length=float(cyl.get("length")))
    if mesh is not None:
        fn = mesh.get("filename")
        if fn.startswith("package://"):
            fn = fn[len("package://"):]
        return Collision(xyz, rpy, "mesh", meshes=load_mesh(os.path.join(base_dir, fn)))
    return None


def load_urdf(path: str) -> Robot:
    tree = ET.parse(path)
    root = tree.getroot()
    base_dir = os.path.dirname(os.path.abspath(path))
    links: List[Link] = []
    for ln in root.findall("link"):
        cols = []
        for c in ln.findall("collision"):
            col = _parse_collision(c, base_dir)
            if col is not None:
                cols.append(col)
        inert = ln.find("inertial")
        ixyz, mass, I = np.zeros(3), 0.0, np.zeros((3, 3))
        if inert is not None:
            o = inert.find("origin")
            ixyz = _floats(o.get("xyz") if o is not None else None, 3)
            m = inert.find("mass")
            mass = float(m.get("value")) if m is not None else 0.0
            it = inert.find("inertia")
            if it is not None:
                g = lambda k: float(it.get(k, 0.0))
                I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")],
                              [g("ixz"), g("iyz"), g("izz")]])
        links.append(Link(ln.get("name"), cols, ixyz, mass, I))
    joints: List[Joint] = []
    for jn in root.findall("joint"):
        o = jn.find("origin")
        ax = jn.find("axis")
        lim = jn.find("limit")
        dyn = jn.find("dynamics")
        lower = upper = None
        if lim is not None:
            lower = float(lim.get("lower")) if lim.get("lower") is not None else None
            upper = float(lim.get("upper")) if lim.get("upper") is not None else None
            # urdfpy fills missing limit bounds with 0.0 only when the tag has neither; the
            # reference treats None as "keep +-1e3" (import_urdf.py:206-214)
        damping = None
        if dyn is not None and dyn.get("damping") is not None:
            damping = float(dyn.get("damping"))
        joints.append(Joint(
            name=jn.get("name"), joint_type=jn.get("type"),
            parent=jn.find("parent").get("link"), child=jn.find("child").get("link"),
            xyz=_floats(o.get("xyz") if o is not None else None, 3),
            rpy=_floats(o.get("rpy") if o is not None else None, 3),
            axis=_floats(ax.get("xyz") if ax is not None else "1 0 0", 3),
            lower=lower, upper=upper, damping=damping))
    return Robot(root.get("name"), links, joints)
