"""rmp2-b200: B200-native engine for the RMP2 control step.

Public surface = the reference's class API (TomGoesGitHub/Riemannian-Motion-Policies):
``kinematics.UrdfForwardKinematic``, ``taskmap.*``, ``rmp.RmpCore`` + v1 leaves, ``rmp2.*`` leaves,
``data_management.Datamanager``.  Compute happens in csrc/ (CUDA, sm_100a) behind the C ABI of
include/rmp2_b200.h; importing this package does not need a GPU, running a step does.
"""
import os

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
URDF_DIR = os.path.join(PACKAGE_DIR, "urdf")

__all__ = ["PACKAGE_DIR", "URDF_DIR"]
