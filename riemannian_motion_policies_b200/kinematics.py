"""Forward kinematics of a URDF robot -- host-side mirror of the reference's ``kinematics.py``.

Same class and call signatures as the reference (kinematics.py:155-270); the arithmetic runs in
the CUDA kernel ``rmp2_fk_kernel`` behind the C ABI (``rmp2_fk``), batched over environments,
with closed-form Jacobian and Jdot*qd instead of TensorFlow autodiff.
"""
import ctypes

import numpy as np
import torch

from . import _native
from ._tensor import current_stream_ptr, like_input, require_cuda, to_device
from .urdf_model import UrdfModel


# ---- small init-time helpers with the reference's names (kinematics.py:22-127) ------------------
def R_x(angle):
    """angle [B,1] -> [B,3,3] (reference: kinematics.py:22-32)."""
    angle = torch.as_tensor(angle, dtype=torch.float32)
    c, s = torch.cos(angle), torch.sin(angle)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([torch.cat([o, z, z], -1), torch.cat([z, c, -s], -1), torch.cat([z, s, c], -1)], dim=-2)


def R_y(angle):
    """reference: kinematics.py:34-44."""
    angle = torch.as_tensor(angle, dtype=torch.float32)
    c, s = torch.cos(angle), torch.sin(angle)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([torch.cat([c, z, s], -1), torch.cat([z, o, z], -1), torch.cat([-s, z, c], -1)], dim=-2)


def R_z(angle):
    """reference: kinematics.py:46-56."""
    angle = torch.as_tensor(angle, dtype=torch.float32)
    c, s = torch.cos(angle), torch.sin(angle)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([torch.cat([c, -s, z], -1), torch.cat([s, c, z], -1), torch.cat([z, z, o], -1)], dim=-2)


def homogenous_transformation(R, t):
    """R [B,3,3], t [B,3] -> T [B,4,4] (reference: kinematics.py:58-71)."""
    R = torch.as_tensor(R, dtype=torch.float32)
    t = torch.as_tensor(t, dtype=torch.float32)
    if R.shape[-2:] != (3, 3) or t.shape[-1] != 3:
        raise ValueError("homogenous_transformation expects R [B,3,3] and t [B,3]")
    T = torch.zeros(R.shape[0], 4, 4, dtype=torch.float32, device=R.device)
    T[:, :3, :3] = R
    T[:, :3, 3] = t
    T[:, 3, 3] = 1.0
    return T


def rotation_matrix_from_rotation_vector(vec, angle):
    """Rodrigues formula, axis used as given (reference: kinematics.py:99-121)."""
    vec = torch.as_tensor(vec, dtype=torch.float32)
    angle = torch.as_tensor(angle, dtype=torch.float32)
    if vec.shape[0] != angle.shape[0]:
        raise ValueError("vec and angle need the same batch size")
    c = torch.cos(angle)[:, None, None]
    s = torch.sin(angle)[:, None, None]
    x, y, z = vec[:, 0], vec[:, 1], vec[:, 2]
    zero = torch.zeros_like(x)
    skew = torch.stack([zero, -z, y, z, zero, -x, -y, x, zero], dim=-1).reshape(-1, 3, 3)
    eye = torch.eye(3, dtype=torch.float32, device=vec.device).expand(vec.shape[0], 3, 3)
    return c * eye + s * skew + (1 - c) * (vec[:, :, None] * vec[:, None, :])


def rotation_matrix_from_rpy(rpy):
    """``R_x(roll) @ R_y(pitch) @ R_z(yaw)`` -- the reference's order (kinematics.py:123-127)."""
    rpy = torch.as_tensor(rpy, dtype=torch.float32)
    return R_x(rpy[:, 0:1]) @ R_y(rpy[:, 1:2]) @ R_z(rpy[:, 2:3])


def euler_from_rotation_matrix(rotation_matrix):
    """xyz Euler angles of rotation matrices [B,3,3] (reference: kinematics.py:74-96)."""
    Rm = torch.as_tensor(rotation_matrix, dtype=torch.float32)
    theta_y = -torch.asin(Rm[:, 2, 0])
    cy = torch.cos(theta_y)
    safe = torch.where(cy.abs() < 1e-6, torch.ones_like(cy), cy)
    theta_z = torch.atan2(Rm[:, 1, 0] / safe, Rm[:, 0, 0] / safe)
    theta_x = torch.atan2(Rm[:, 2, 1] / safe, Rm[:, 2, 2] / safe)
    return torch.stack((theta_x, theta_y, theta_z), dim=-1)


class UrdfForwardKinematic:
    """A generic kinematic class for any URDF file (reference: kinematics.py:155-270).

    ``forward`` / ``differentiate`` accept ``q`` of shape [1, n] like the reference, and also
    [B, n] for B environments at once.
    """

    def __init__(self, urdf_filepath, order):
        self.filepath = urdf_filepath
        self.order = list(order)
        self.n_joints = len(self.order)
        self._handle = None
        self._build()

    def _build(self):
        """Parse the URDF and hand the constant tables to the native library
        (reference: kinematics.py:163-210)."""
        model = UrdfModel(self.filepath)
        self.model = model
        self.frame_names = model.frame_names
        self._name_to_idx = {name: i for i, name in enumerate(self.frame_names)}
        # column of q per frame; frames whose joint is not in `order` read a constant 0
        # (reference: kinematics.py:197, 218-219)
        self._q_reordering = np.array(
            [self.order.index(name) if name in self.order else -1 for name in self.frame_names], dtype=np.int32)
        self.T_constant = model.constant_transforms()
        self.axis = model.axes()
        self.joint_type = model.type_codes()
        self.parent = model.parents()
        self.has_collision = [f.has_collision for f in model.frames]
        handle = ctypes.c_void_p()
        T = np.ascontiguousarray(self.T_constant.reshape(-1, 16))
        _native.check(_native.lib().rmp2_robot_create(
            T.ctypes.data, self.axis.ctypes.data, self.joint_type.ctypes.data, self.parent.ctypes.data,
            self._q_reordering.ctypes.data, len(self.frame_names), self.n_joints, ctypes.byref(handle)))
        self._handle = handle

    def __del__(self):
        try:
            if self._handle is not None:
                _native.lib().rmp2_robot_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def frame_index(self, frame):
        if isinstance(frame, bytes):
            frame = frame.decode()
        if hasattr(frame, "numpy") and not isinstance(frame, str):
            frame = frame.numpy()
            frame = frame.decode() if isinstance(frame, bytes) else str(frame)
        if frame not in self._name_to_idx:
            raise KeyError(f"unknown frame {frame!r}; frames: {self.frame_names}")
        return self._name_to_idx[frame]

    def _run(self, q, qd, frame, derivatives):
        dev = require_cuda()
        qt = to_device(q, dev).reshape(-1, self.n_joints)
        B = qt.shape[0]
        x = torch.empty(B, 16, device=dev, dtype=torch.float32)
        if derivatives:
            qdt = to_device(qd, dev).reshape(-1, self.n_joints)
            if qdt.shape[0] != B:
                raise ValueError("q and qd need the same batch size")
            xd = torch.empty(B, 16, device=dev, dtype=torch.float32)
            J = torch.empty(B, 16, self.n_joints, device=dev, dtype=torch.float32)
            c = torch.empty(B, 16, device=dev, dtype=torch.float32)
            ptrs = (qdt.data_ptr(), x.data_ptr(), xd.data_ptr(), J.data_ptr(), c.data_ptr())
        else:
            xd = J = c = None
            ptrs = (None, x.data_ptr(), None, None, None)
        _native.check(_native.lib().rmp2_fk(self._handle, self.frame_index(frame), B, qt.data_ptr(), *ptrs,
                                            current_stream_ptr(dev)))
        return x, xd, J, c

    def forward(self, q, frame):
        """q [1,n] (or [B,n]) -> T [1,4,4] (or [B,4,4])  (reference: kinematics.py:212-247)."""
        x, _, _, _ = self._run(q, None, frame, derivatives=False)
        return like_input(x.reshape(-1, 4, 4), q)

    __call__ = forward      # reference: taskmap.py:28 calls ``self.fkine(q, self.frame)``

    def differentiate(self, q, qd, frame):
        """-> x [B,16], xd [B,16], J [B,16,n], c [B,16]  (reference: kinematics.py:250-270)."""
        x, xd, J, c = self._run(q, qd, frame, derivatives=True)
        return tuple(like_input(t, q) for t in (x, xd, J, c))
