"""URDF -> frame table: product reader and oracle reader against the reference's own parser.

tests/golden/urdf_frames.json was produced by importing the reference's helper/urdf_parsing.py on the
reference's URDF files (tests/golden/make_golden.py).  The URDFs shipped in this repo are kinematic-only
re-statements; they must parse to the same frame table."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import rmp_oracle as O
from riemannian_motion_policies_b200 import scenarios as S
from riemannian_motion_policies_b200.urdf_model import UrdfModel

with open(os.path.join(GOLDEN, "urdf_frames.json")) as fh:
    REF = json.load(fh)

FILES = {"panda": S.PANDA_URDF, "panda_wo_tool": S.PANDA_WO_TOOL_URDF, "two_joint": S.TWO_JOINT_URDF,
         "gantry": S.GANTRY_URDF}
REFERENCE_FILES = {"panda": "/root/reference/urdf/franka_panda/panda.urdf",
                   "panda_wo_tool": "/root/reference/urdf/franka_panda/panda_wo_tool.urdf",
                   "two_joint": "/root/reference/urdf/TwoJointRobot_wo_fixedJoints.urdf"}


def _check_against_reference(frames_ref, names, types, rpy, xyz, axis, col, paths):
    assert names == [f["name"] for f in frames_ref]
    assert types == [f["joint_type"] for f in frames_ref]
    assert col == [f["has_collision"] for f in frames_ref]
    np.testing.assert_array_equal(np.array(rpy), np.array([f["rpy"] for f in frames_ref]))
    np.testing.assert_array_equal(np.array(xyz), np.array([f["xyz"] for f in frames_ref]))
    np.testing.assert_array_equal(np.array(axis), np.array([f["axis"] for f in frames_ref]))
    assert paths == [f["path"] for f in frames_ref]


@pytest.mark.parametrize("key", sorted(FILES))
def test_oracle_reader_matches_reference_parser(key):
    fr = O.urdf_frames(FILES[key])
    _check_against_reference(REF[key], [f["name"] for f in fr], [f["joint_type"] for f in fr], [f["rpy"] for f in fr],
                             [f["xyz"] for f in fr], [f["axis"] for f in fr], [f["has_collision"] for f in fr],
                             [f["path"] for f in fr])


@pytest.mark.parametrize("key", sorted(FILES))
def test_product_reader_matches_reference_parser(key):
    m = UrdfModel(FILES[key])
    names = m.frame_names
    paths = [[names[i] for i in f.chain] for f in m.frames]
    _check_against_reference(REF[key], names, [f.joint_type for f in m.frames], [f.rpy for f in m.frames],
                             [f.xyz for f in m.frames], [f.axis for f in m.frames], [f.has_collision for f in m.frames], paths)


@pytest.mark.parametrize("key", sorted(REFERENCE_FILES))
def test_readers_on_the_reference_urdfs_when_present(key):
    """In the build container the reference's full URDFs (meshes, inertias) are read directly."""
    path = REFERENCE_FILES[key]
    if not os.path.exists(path):
        pytest.skip("reference checkout not present (GPU box)")
    m = UrdfModel(path)
    assert m.frame_names == [f["name"] for f in REF[key]]
    assert [f.has_collision for f in m.frames] == [f["has_collision"] for f in REF[key]]
    fr = O.urdf_frames(path)
    assert [f["name"] for f in fr] == m.frame_names


def test_constant_transforms_follow_the_reference_rpy_order():
    """R_x(r) @ R_y(p) @ R_z(y), float32 (reference: kinematics.py:123-127, 200-203)."""
    import torch
    m = UrdfModel(S.PANDA_URDF)
    T = m.constant_transforms()
    fk = O.UrdfForwardKinematic(S.PANDA_URDF, S.PANDA_ORDER_9)
    np.testing.assert_allclose(T, fk.T_constant.numpy(), atol=1e-7)
    assert T.dtype == np.float32
    # multi-axis rpy: the reference order differs from URDF's R_z R_y R_x
    rpy = torch.tensor([[0.3, -0.4, 0.5]])
    R = O.rotation_matrix_from_rpy(rpy)[0].numpy()
    from scipy.spatial.transform import Rotation
    expected = (Rotation.from_euler("x", 0.3) * Rotation.from_euler("y", -0.4) * Rotation.from_euler("z", 0.5)).as_matrix()
    np.testing.assert_allclose(R, expected, atol=1e-6)


def test_unsupported_joint_type_is_rejected(tmp_path):
    p = tmp_path / "bad.urdf"
    p.write_text('<robot name="r"><link name="a"/><link name="b"/>'
                 '<joint name="j" type="floating"><origin rpy="0 0 0" xyz="0 0 0"/><parent link="a"/><child link="b"/></joint></robot>')
    with pytest.raises(NotImplementedError):
        UrdfModel(str(p))
