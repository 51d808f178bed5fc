"""Pins the oracle's kinematics the way the reference's own tests pin the reference.

Re-expression of reference tests/test_kinematic_forwards.py:16-137 (SciPy ground truth, same
tolerances).  Where the reference uses PyBullet as ground truth (not installed), an independent
float64 SciPy forward kinematics and finite differences stand in."""
import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

from oracle import rmp_oracle as O
from riemannian_motion_policies_b200 import scenarios as S
from riemannian_motion_policies_b200.urdf_model import UrdfModel


def test_R():
    """reference: tests/test_kinematic_forwards.py:16-37"""
    rng = np.random.RandomState(0)
    for fn, vec in zip([O.R_x, O.R_y, O.R_z], np.eye(3)):
        for _ in range(10):
            bs = rng.randint(1, 9)
            angle = rng.uniform(0, 2 * np.pi, size=bs)
            truth = np.array([Rotation.from_rotvec(a * vec).as_matrix() for a in angle]).reshape(bs, 3, 3)
            got = fn(torch.tensor(angle.astype(np.float32)).reshape(bs, 1)).numpy()
            assert np.max(np.abs(truth - got)) <= 1e-6
            assert got.shape == (bs, 3, 3)


def test_homogenous_transformation():
    """reference: tests/test_kinematic_forwards.py:39-59"""
    rng = np.random.RandomState(1)
    for _ in range(10):
        bs = rng.randint(1, 9)
        R = rng.uniform(0, 99, size=(bs, 3, 3))
        t = rng.uniform(0, 99, size=(bs, 3, 1))
        T = np.concatenate([np.concatenate([R, t], -1), np.broadcast_to([[0, 0, 0, 1]], (bs, 1, 4))], -2)
        got = O.homogenous_transformation(torch.tensor(R, dtype=torch.float32), torch.tensor(t.reshape(bs, 3), dtype=torch.float32)).numpy()
        assert np.max(np.abs(T - got)) <= 1e-5
        assert got.shape == (bs, 4, 4)


def test_rotation_matrix_from_rotation_vector():
    """reference: tests/test_kinematic_forwards.py:61-85"""
    rng = np.random.RandomState(2)
    for _ in range(10):
        bs = rng.randint(1, 9)
        vec = rng.uniform(size=(bs, 3))
        vec /= np.linalg.norm(vec, axis=-1, keepdims=True)
        angle = rng.uniform(0, 2 * np.pi, size=bs)
        truth = np.array([Rotation.from_rotvec(a * v).as_matrix() for v, a in zip(vec, angle)])
        got = O.rotation_matrix_from_rotation_vector(torch.tensor(vec, dtype=torch.float32), torch.tensor(angle, dtype=torch.float32)).numpy()
        assert np.max(np.abs(truth - got)) <= 1e-6


def test_euler_from_rotation_matrix():
    """reference: tests/test_kinematic_forwards.py:87-106 (round trip through R, 1e-4)"""
    rng = np.random.RandomState(3)
    for _ in range(100):
        bs = rng.randint(1, 9)
        eulers = rng.uniform(0, 2 * np.pi, size=(bs, 3))
        R = np.array([Rotation.from_euler("xyz", e).as_matrix() for e in eulers], dtype=np.float32)
        got = O.euler_from_rotation_matrix(torch.tensor(R)).numpy()
        R2 = np.array([Rotation.from_euler("xyz", e).as_matrix() for e in got], dtype=np.float32)
        assert np.max(np.abs(R - R2)) < 1e-4
        assert got.shape == (bs, 3)


def _scipy_fk(model, order, q):
    """Independent float64 FK with the URDF convention (extrinsic xyz = R_z R_y R_x); stands in for
    PyBullet's getLinkState in reference tests/test_kinematic_forwards.py:108-137.  The shipped
    robots only have single-axis rpy, for which both conventions coincide."""
    world = []
    for f in model.frames:
        T = np.eye(4)
        T[:3, :3] = Rotation.from_euler("xyz", f.rpy).as_matrix()
        T[:3, 3] = f.xyz
        qi = q[order.index(f.name)] if f.name in order else 0.0
        V = np.eye(4)
        if f.joint_type == "revolute":
            V[:3, :3] = Rotation.from_rotvec(qi * np.array(f.axis)).as_matrix()
        elif f.joint_type == "prismatic":
            V[:3, 3] = qi * np.array(f.axis)
        parent = world[f.parent] if f.parent >= 0 else np.eye(4)
        world.append(parent @ T @ V)
    return world


@pytest.mark.parametrize("urdf,order", [(S.PANDA_URDF, S.PANDA_ORDER_9), (S.PANDA_WO_TOOL_URDF, S.PANDA_ORDER_7),
                                        (S.TWO_JOINT_URDF, S.TWO_JOINT_ORDER)])
def test_fk_of_every_frame(urdf, order):
    """reference: tests/test_kinematic_forwards.py:108-137, tolerance 1e-6 in float32."""
    fk = O.UrdfForwardKinematic(urdf, order)
    model = UrdfModel(urdf)
    n = len(order)
    rng = np.random.RandomState(4)
    lo, hi = (S.PANDA_Q_LOW[:n], S.PANDA_Q_HIGH[:n]) if n > 2 else (-np.pi * np.ones(2), np.pi * np.ones(2))
    for _ in range(25):
        q = rng.uniform(lo, hi)
        truth = _scipy_fk(model, order, q)
        for i, name in enumerate(fk.frame_names):
            T = fk.forward(torch.tensor(q, dtype=torch.float32)[None], name)[0].numpy()
            assert np.max(np.abs(T - truth[i])) < 2e-6, name


def test_fk_derivatives_against_finite_differences():
    """Stands in for reference tests/test_kinematic_differentiability.py:24-74 (PyBullet Jacobian):
    J, xd = J qd and c = d(J qd)/dq . qd of the float64 oracle against central differences."""
    fk = O.UrdfForwardKinematic(S.PANDA_URDF, S.PANDA_ORDER_9, dtype=torch.float64)
    rng = np.random.RandomState(5)
    h = 1e-6
    for frame in ("panda_joint4", "panda_finger_joint2", "panda_grasptarget_hand"):
        q = torch.tensor(rng.uniform(S.PANDA_Q_LOW, S.PANDA_Q_HIGH))
        qd = torch.tensor(rng.uniform(-1, 1, size=9))
        x, xd, J, c = fk.differentiate(q[None], qd[None], frame)
        f = lambda qq: fk.forward(qq[None], frame).reshape(-1)
        J_fd = torch.stack([(f(q + h * e) - f(q - h * e)) / (2 * h) for e in torch.eye(9, dtype=torch.float64)], dim=1)
        assert (J[0] - J_fd).abs().max() < 1e-8
        assert (xd[0] - J[0] @ qd).abs().max() < 1e-12
        vel = lambda qq: fk.differentiate(qq[None], qd[None], frame)[1][0]
        c_fd = (vel(q + h * qd) - vel(q - h * qd)) / (2 * h)
        assert (c[0] - c_fd).abs().max() < 1e-7


def test_distance_map_has_gradient_only_through_the_frame_origin():
    """taskmap.py:120-138: value = |p_link - p_obs|, derivative as if the link point translated with
    the frame origin; closed form J = n^T J_pos, c = n.c_pos + (|v|^2 - (n.v)^2)/d."""
    fk = O.UrdfForwardKinematic(S.PANDA_URDF, S.PANDA_ORDER_9, dtype=torch.float64)
    rng = np.random.RandomState(6)
    q = torch.tensor(rng.uniform(S.PANDA_Q_LOW, S.PANDA_Q_HIGH))
    qd = torch.tensor(rng.uniform(-1, 1, size=9))
    link = torch.tensor(rng.uniform(-0.5, 0.5, size=(5, 3)))
    obst = torch.tensor(rng.uniform(-0.5, 0.5, size=(5, 3)))
    tm = O.chain_taskmaps([O.TaskmapByForwardKinematic(fk, "panda_joint6"), O.TaskmapJointFrame4x4ToDistance(link, obst)])
    x, xd, J, c = tm.differentiate(q[None], qd[None])
    T, Td, JT, cT = fk.differentiate(q[None], qd[None], "panda_joint6")
    rows = [3, 7, 11]
    v, ck, Jk = Td[0][rows], cT[0][rows], JT[0][rows]
    r = link - obst
    d = r.norm(dim=1)
    n = r / d[:, None]
    assert (x[:, 0] - d).abs().max() < 1e-12
    assert (xd[:, 0] - n @ v).abs().max() < 1e-12
    assert (J[:, 0] - n @ Jk).abs().max() < 1e-12
    assert (c[:, 0] - (n @ ck + (v @ v - (n @ v) ** 2) / d)).abs().max() < 1e-11
