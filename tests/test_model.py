"""Model compiler facts that SURVEY.md section 8 pins (body / dof / contact counts, tree shape, post-processing)."""
import numpy as np
import pytest

from ppr_diffphys_b200 import load_robot
from ppr_diffphys_b200.model import JOINT_COMPOUND, JOINT_FREE, JOINT_REVOLUTE, RobotModel, mesh_mass_inertia

EXPECT = {
    "laikago": dict(nb=13, nq=19, nqd=18, nc=3848, parents=[-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11],
                    kp=220.0, kd=2.0, attach=(16000.0, 200.0), jt=JOINT_REVOLUTE),
    "human": dict(nb=19, nq=61, nqd=60, nc=152, parents=[-1, 0, 1, 2, 3, 3, 5, 6, 7, 3, 9, 10, 11, 0, 13, 14, 0, 16, 17],
                  kp=660.0, kd=5.0, attach=(8000.0, 200.0), jt=JOINT_COMPOUND),
    "quad": dict(nb=26, nq=82, nqd=81, nc=208,
                 parents=[-1, 0, 1, 2, 3, 3, 5, 6, 7, 3, 9, 10, 11, 0, 13, 14, 15, 16, 0, 18, 19, 20, 0, 22, 23, 24],
                 kp=660.0, kd=5.0, attach=(8000.0, 200.0), jt=JOINT_COMPOUND),
}


@pytest.mark.parametrize("robot", sorted(EXPECT))
def test_compiled_robot_matches_survey(robot):
    m, e = load_robot(robot), EXPECT[robot]
    assert (m.nb, m.nq, m.nqd, m.nc) == (e["nb"], e["nq"], e["nqd"], e["nc"])
    assert m.joint_parent.tolist() == e["parents"]
    assert m.joint_type[0] == JOINT_FREE and set(m.joint_type[1:].tolist()) == {e["jt"]}
    assert np.all(m.joint_target_ke[:6] == 0) and np.all(m.joint_target_ke[6:] == e["kp"])
    assert np.all(m.joint_target_kd[:6] == 0) and np.all(m.joint_target_kd[6:] == e["kd"])
    assert (m.joint_attach_ke, m.joint_attach_kd) == e["attach"]
    assert np.allclose(m.shape_materials, [[1e4, 0.0, 1e2, 1.0]] * len(m.shape_materials))
    assert np.all(m.joint_limit_ke == 0) and np.all(m.joint_limit_kd == 0)
    assert np.all(np.linalg.eigvalsh(m.norm_body_inertia.astype(np.float64)) > 0)
    assert np.all(m.body_mass > 0)
    # joints in topological order (eval_fk reads the parent from its own output)
    assert all(p < i for i, p in enumerate(m.joint_parent))
    assert np.allclose(np.linalg.norm(m.joint_X_p[:, 3:], axis=1), 1.0, atol=1e-6)


def test_box_robots_mass_clamp_and_feet_scaling():
    m = load_robot("human")
    assert m.body_mass.min() >= 1.0 and m.body_mass.max() <= 5.0
    feet = [i for i, n in enumerate(m.body_names) if n in ("link_24_mixamorig:RightFoot_Y", "link_19_mixamorig:LeftFoot_Y")]
    assert len(feet) == 2
    for f in feet:  # feet boxes doubled (dp_model.py:169-177): contact corners of a foot span twice a plain box
        pts = m.contact_point[m.contact_body == f]
        assert pts.shape == (8, 3)


def test_mesh_mass_of_unit_cube():
    V = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], dtype=float)
    F = np.array([[0, 2, 3], [0, 3, 1], [4, 5, 7], [4, 7, 6], [0, 1, 5], [0, 5, 4], [2, 6, 7], [2, 7, 3],
                  [0, 4, 6], [0, 6, 2], [1, 3, 7], [1, 7, 5]])
    vol, I, com = mesh_mass_inertia(V, F[:, ::-1])  # outward-facing winding
    assert abs(vol - 1.0) < 1e-12 and np.allclose(com, 0.5)
    assert np.allclose(I, np.eye(3) / 6.0, atol=1e-12)


def test_npz_roundtrip(tmp_path):
    m = load_robot("laikago")
    p = str(tmp_path / "m.npz")
    m.save(p)
    m2 = RobotModel.load(p)
    assert m2.nc == m.nc and np.array_equal(m2.contact_point, m.contact_point) and m2.body_names == m.body_names
