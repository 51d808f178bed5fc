"""Stage-level parity: FK / Jacobian / Jdot*qd of every frame and every leaf policy, CUDA vs oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from gpu_common import product_fkine
from oracle import harness as H
from oracle import rmp_oracle as O
from riemannian_motion_policies_b200 import scenarios as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ns(native_lib):
    return S.product_namespace()


@pytest.mark.parametrize("n", [2, 7, 9])
def test_fk_differentiate_every_frame(ns, n):
    """x, xd, J, c of all 16 entries of T for every frame against the float64 autodiff oracle
    (golden fixture fk_n*.npz).  Tolerances: the reference's own 1e-6 for FK and J
    (tests/test_kinematic_forwards.py:137, tests/test_taskmaps.py:74), scaled for velocities."""
    g = np.load(os.path.join(GOLDEN, f"fk_n{n}.npz"))
    fk = product_fkine(ns, n)
    assert list(g["frame_names"]) == fk.frame_names
    q, qd = torch.as_tensor(g["q"]).cuda(), torch.as_tensor(g["qd"]).cuda()
    for fi, frame in enumerate(fk.frame_names):
        x, xd, J, c = (t.cpu().numpy() for t in fk.differentiate(q, qd, frame))
        assert x.shape == (q.shape[0], 16) and J.shape == (q.shape[0], 16, n)
        np.testing.assert_allclose(x, g[f"x_{fi}"], atol=2e-6, err_msg=frame)
        np.testing.assert_allclose(J, g[f"J_{fi}"], atol=2e-6, err_msg=frame)
        np.testing.assert_allclose(xd, g[f"xd_{fi}"], atol=5e-6, err_msg=frame)
        np.testing.assert_allclose(c, g[f"c_{fi}"], atol=2e-5, err_msg=frame)
        T = fk.forward(q, frame).cpu().numpy()
        np.testing.assert_array_equal(T.reshape(-1, 16), x)


def test_fk_general_axes_prismatic_xyz_and_branches(ns):
    """The synthetic gantry arm (urdf/make_urdf.py): revolute joints about general unit axes (the full Rodrigues
    branch of chain_advance, reference kinematics.py:99-121), prismatic joints along x, y and z, constant rotations
    with all three rpy components (R_x R_y R_z order, kinematics.py:123-127), a tree that branches three times and a
    scrambled joint order -- x, xd, J, c of every frame against the float64 autodiff oracle, computed live."""
    fk = ns.UrdfForwardKinematic(S.GANTRY_URDF, S.GANTRY_ORDER)
    ofk = H.make_fkine(9, torch.float64, robot="gantry")
    assert fk.frame_names == ofk.frame_names
    q, qd, _ = S.sample_gantry_state(6, seed=31)
    qd = (3 * qd).astype(np.float32)
    tq, tqd = torch.as_tensor(q).cuda(), torch.as_tensor(qd).cuda()
    for frame in fk.frame_names:
        x, xd, J, c = (t.cpu().numpy() for t in fk.differentiate(tq, tqd, frame))
        for b in range(q.shape[0]):
            xo, xdo, Jo, co = (t[0].numpy() for t in ofk.differentiate(torch.as_tensor(q[b:b + 1]).double(),
                                                                         torch.as_tensor(qd[b:b + 1]).double(), frame))
            np.testing.assert_allclose(x[b], xo, atol=2e-6, err_msg=frame)
            np.testing.assert_allclose(J[b], Jo, atol=2e-6, err_msg=frame)
            np.testing.assert_allclose(xd[b], xdo, atol=5e-6, err_msg=frame)
            np.testing.assert_allclose(c[b], co, atol=3e-5, err_msg=frame)


def test_fk_reference_call_shapes(ns):
    """q [1,n] numpy in -> [1,4,4] host tensor out, callable alias (taskmap.py:28)."""
    fk = product_fkine(ns, 9)
    T = fk.forward(np.array([S.PANDA_Q_READY], dtype=np.float32), "panda_grasptarget_hand")
    assert tuple(T.shape) == (1, 4, 4) and not T.is_cuda
    assert torch.equal(T, fk(np.array([S.PANDA_Q_READY], dtype=np.float32), "panda_grasptarget_hand"))
    with pytest.raises(KeyError):
        fk.forward(np.zeros((1, 9), np.float32), "no_such_frame")


def test_euler_taskmap_chain(ns):
    """chain [FK, TaskmapFrom4x4ToEuler] (reference: tests/test_taskmaps.py:18-76 checks this chain's J against
    PyBullet's angular Jacobian at 1e-3): x, xd, J, c against the float64 autodiff oracle."""
    fk = product_fkine(ns, 7)
    ofk = H.make_fkine(7, torch.float64)
    tm = ns.chain_taskmaps([ns.TaskmapByForwardKinematic(fk, "panda_joint6"), ns.TaskmapFrom4x4ToEuler()])
    otm = O.chain_taskmaps([O.TaskmapByForwardKinematic(ofk, "panda_joint6"), O.TaskmapFrom4x4ToEuler()])
    q, qd, _ = S.sample_panda_state(5, 7, seed=9)
    for b in range(5):
        x, xd, J, c = (t.numpy() for t in tm.differentiate(q[b:b + 1], qd[b:b + 1]))
        xo, xdo, Jo, co = (t.numpy() for t in otm.differentiate(torch.as_tensor(q[b:b + 1]).double(), torch.as_tensor(qd[b:b + 1]).double()))
        np.testing.assert_allclose(x, xo, atol=1e-5)
        np.testing.assert_allclose(J, Jo, atol=1e-4)
        np.testing.assert_allclose(xd, xdo, atol=1e-4)
        np.testing.assert_allclose(c, co, atol=1e-3)


def _leaf_pairs(ns, ons):
    lim_lo, lim_hi = S.PANDA_Q_LOW[:7], S.PANDA_Q_HIGH[:7]
    mk = lambda m: (
        ("TargetPolicy3", m.TargetPolicy(0.1, 1, 0.1, [0.4, -0.1, 0.6], m.IdentityTaskmap()), 3),
        ("TargetPolicy7", m.TargetPolicy(0.1, 0.5, 0.1, list(S.CSPACE_GOAL_9[:7]), m.IdentityTaskmap()), 7),
        ("TargetAttractor", m.TargetAttractor([0.4, -0.1, 0.6], 0.3, 0.6, 0.075, 0.05, 0.03, 1, 0.5, 1., 0.02, m.IdentityTaskmap()), 3),
        ("ConfigurationSpaceBiasing", m.ConfigurationSpaceBiasing(0.01, 0.1, S.NULLSPACE_Q0_9[:7], 'b', w=0.05), 7),
        ("JointLimitAvoidance", m.JointLimitAvoidance(lim_lo, lim_hi, 0.3, 1), 7),
        ("JointVelocityCap", m.JointVelocityCap(0.5, 0.15, 5.0, 0.05), 7),
        ("JointDamping", m.JointDamping(1, 0.005, 0.3), 7),
        ("CSpaceBiasing", m.CSpaceBiasing(S.CSPACE_GOAL_9[:7], 0.005, 1, 2, 0.5, 0.0001), 7),
        ("ObstacleAvoidance", m.ObstacleAvoidance(0., 50, 0.04, 0.01, 0.01, 800, 0.01, 0.5, 1, 0.02, 0.001, m.IdentityTaskmap(), 'o'), 1),
    )
    return list(zip(mk(ns), mk(ons)))


def test_v1_collision_avoidance_leaf(ns):
    """CollisionAvoidance.evaluate(x, xd) with its own d / vec data (rmp.py:264-315)."""
    rng = np.random.RandomState(5)
    K = 50
    d = rng.uniform(0.02, 1.5, size=K).astype(np.float32)
    vec = rng.normal(size=(K, 3))
    vec = (vec / np.linalg.norm(vec, axis=1, keepdims=True)).astype(np.float32)
    x = rng.uniform(-1, 1, size=(K, 3)).astype(np.float32)
    xd = rng.uniform(-0.5, 0.5, size=(K, 3)).astype(np.float32)
    args = dict(eta_rep=0.1 * np.e, nu_rep=0.3, eta_damp=1, nu_damp=0.3, r=1.1, c=1e5)
    f, A = ns.CollisionAvoidance(d, vec, taskmap=ns.IdentityTaskmap(), **args).evaluate(x, xd)
    fo, Ao = O.CollisionAvoidance(torch.as_tensor(d).double(), torch.as_tensor(vec).double(), taskmap=None, **args).evaluate(
        torch.as_tensor(x).double(), torch.as_tensor(xd).double())
    np.testing.assert_allclose(f.numpy(), fo.numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(A.numpy(), Ao.numpy(), rtol=2e-5, atol=2e-6)


def test_every_leaf_policy(ns):
    """leaf.evaluate(x, xd) -> (xdd, M) for every leaf class against the oracle in float64
    (single row at a time where the reference's norm is a whole-tensor norm, rmp.py:243)."""
    ons = H.namespace(torch.float64)
    rng = np.random.RandomState(0)
    for (name, leaf, m), (_, oleaf, _) in _leaf_pairs(ns, ons):
        K = 64
        if m == 1:
            x = rng.uniform(0.02, 0.7, size=(K, 1))
            xd = rng.uniform(-0.05, 0.05, size=(K, 1))
        elif m == 3:
            x = rng.uniform(-0.8, 0.8, size=(K, 3))
            xd = rng.uniform(-0.5, 0.5, size=(K, 3))
        else:
            x = rng.uniform(S.PANDA_Q_LOW[:7], S.PANDA_Q_HIGH[:7], size=(K, 7))
            xd = rng.uniform(-0.6, 0.6, size=(K, 7))
            if name == "JointVelocityCap":
                # keep clear of the metric's poles |qd| = v_max - 2*region = 0.2 and of the clipped zone
                # |qd| >= v_max, where 1 - ratio^2 ~ 1e-5 amplifies float32 rounding
                mag = np.where(rng.rand(K, 7) < 0.5, rng.uniform(0.0, 0.17, size=(K, 7)), rng.uniform(0.25, 0.48, size=(K, 7)))
                xd = mag * np.sign(xd)
        x32, xd32 = x.astype(np.float32), xd.astype(np.float32)
        xdd, M = leaf.evaluate(x32, xd32)
        assert tuple(xdd.shape) == (K, m) and tuple(M.shape) == (K, m, m)
        for k in range(K):
            a, A = oleaf.evaluate(torch.as_tensor(x32[k:k + 1]).double(), torch.as_tensor(xd32[k:k + 1]).double())
            np.testing.assert_allclose(xdd[k].numpy(), a[0].numpy(), rtol=2e-5, atol=2e-6, err_msg=name)
            np.testing.assert_allclose(M[k].numpy(), A[0].numpy(), rtol=2e-5, atol=2e-6, err_msg=name)
