"""Task maps -- host-side mirror of the reference's ``taskmap.py`` (same names and signatures).

A chain built with ``chain_taskmaps([...])`` remembers its stages, so ``RmpCore`` can recognise
the closed set of chains the CUDA step kernel implements and compile the
tree; the kernel then evaluates the whole chain analytically.  The stand-alone ``forward`` /
``differentiate`` methods below exist for API parity: FK runs in the CUDA FK kernel, and the small
closed-form maps on top of it run as torch ops on the CUDA tensors it returns.
"""
import torch

from ._tensor import like_input, require_cuda, to_device, unwrap


class Taskmap:
    """reference: taskmap.py:6-11."""

    def forward(self, q):
        raise NotImplementedError

    def differentiate(self, q, qd):
        raise NotImplementedError


class IdentityTaskmap(Taskmap):
    """x = q (reference: taskmap.py:13-20)."""

    def forward(self, q):
        return q

    def differentiate(self, q, qd):
        dev = require_cuda()
        qt, qdt = to_device(q, dev), to_device(qd, dev)
        J = torch.eye(qt.shape[-1], device=dev).expand(qt.shape[0], -1, -1).contiguous()
        out = (qt, qdt, J, torch.zeros_like(qt))
        return tuple(like_input(t, q) for t in out)


class TaskmapByForwardKinematic(Taskmap):
    """reference: taskmap.py:22-31."""

    def __init__(self, fkine, frame):
        self.fkine = fkine
        self.frame = frame.decode() if isinstance(frame, bytes) else str(frame)
        fkine.frame_index(self.frame)          # KeyError now rather than at the first step

    def forward(self, q):
        return self.fkine(q, self.frame)

    def differentiate(self, q, qd):
        return self.fkine.differentiate(q, qd, self.frame)


class TaskmapByFunction(Taskmap):
    """reference: taskmap.py:33-42.  ``stages`` is set by chain_taskmaps (None for user functions,
    which the tree compiler rejects: there is no generic-autodiff fallback)."""

    def __init__(self, forward_fn, differentiate_fn, stages=None):
        self.forward_fn = forward_fn
        self.differentiate_fn = differentiate_fn
        self.stages = stages

    def forward(self, q):
        return self.forward_fn(q)

    def differentiate(self, q, qd):
        return self.differentiate_fn(q, qd)


class TaskmapFrom4x4ToPosition(Taskmap):
    """pos = T[:3, 3] (reference: taskmap.py:45-54)."""

    def forward(self, input):
        t = to_device(input)
        return like_input(t.reshape(-1, 4, 4)[:, :3, 3].contiguous(), input)

    def differentiate(self, q, qd):
        dev = require_cuda()
        qt, qdt = to_device(q, dev).reshape(-1, 16), to_device(qd, dev).reshape(-1, 16)
        rows = torch.tensor([3, 7, 11], device=dev)
        J = torch.zeros(qt.shape[0], 3, 16, device=dev)
        J[:, torch.arange(3, device=dev), rows] = 1.0
        out = (qt[:, rows], qdt[:, rows], J, torch.zeros(qt.shape[0], 3, device=dev))
        return tuple(like_input(t, q) for t in out)


class _DistanceBase(Taskmap):
    """Closed form of the reference's autodiff through ``norm(p_joint + stop_grad(p_link - p_joint) - p_obs)``."""

    def _points(self, dev):
        raise NotImplementedError

    def forward(self, input):
        link, obst = self._points(require_cuda())
        d = torch.linalg.norm(link - obst, dim=-1)[:, None]
        return like_input(d, input)

    def differentiate(self, q, qd):
        dev = require_cuda()
        link, obst = self._points(dev)
        K = link.shape[0]
        qdt = to_device(qd, dev).reshape(-1, 16)
        r = link - obst
        d = torch.linalg.norm(r, dim=-1)
        nhat = r / d[:, None]
        v = qdt[:, [3, 7, 11]].expand(K, 3)
        xdot = (nhat * v).sum(-1)
        J = torch.zeros(K, 1, 16, device=dev)
        J[:, 0, [3, 7, 11]] = nhat
        c = ((v * v).sum(-1) - xdot * xdot) / d
        out = (d[:, None], xdot[:, None], J, c[:, None])
        return tuple(like_input(t, q) for t in out)


class TaskmapJointFrame4x4ToDistance(_DistanceBase):
    """Distance of K closest-point pairs attached to one frame (reference: taskmap.py:115-138).
    The two arguments may be tensors/arrays [K,3] or variables that are updated in place
    (``Datamanager`` entries, data_management.py:8-17); they are read again at every step."""

    def __init__(self, pos_on_link_in_base_frame, pos_on_obstacle_in_base_frame):
        self.pos_on_link_in_base_frame = pos_on_link_in_base_frame
        self.pos_on_obstacle_in_base_frame = pos_on_obstacle_in_base_frame

    def _points(self, dev):
        link = to_device(self.pos_on_link_in_base_frame, dev).reshape(-1, 3)
        obst = to_device(self.pos_on_obstacle_in_base_frame, dev).reshape(-1, 3)
        if link.shape != obst.shape:
            raise ValueError("pos_on_link_in_base_frame and pos_on_obstacle_in_base_frame need the same shape")
        return link, obst

    def current_pairs(self):
        """[K,8] pair rows (pos_on_link, pos_on_obstacle, 0, 0) read from the holders right now."""
        link = unwrap(self.pos_on_link_in_base_frame)
        obst = unwrap(self.pos_on_obstacle_in_base_frame)
        link = torch.as_tensor(link, dtype=torch.float32).reshape(-1, 3).cpu()
        obst = torch.as_tensor(obst, dtype=torch.float32).reshape(-1, 3).cpu()
        if link.shape != obst.shape:
            raise ValueError("pos_on_link_in_base_frame and pos_on_obstacle_in_base_frame need the same shape")
        return torch.cat([link, obst, torch.zeros(link.shape[0], 2)], dim=1)


class TaskmapJointFrame4x4ToSphereDistance(Taskmap):
    """Additive to the reference: the K pairs of this frame are the O sphere obstacles passed to
    ``RmpCore.evaluate(..., spheres=[B,O,4])``: pos_on_link = frame origin, pos_on_obstacle = the
    closest point of the sphere surface.  Only meaningful inside a compiled tree."""

    def forward(self, input):
        raise NotImplementedError("sphere distances are evaluated inside RmpCore.evaluate(..., spheres=...)")

    differentiate = forward


class TaskmapFrom4x4ToEuler(Taskmap):
    """xyz Euler angles of the rotation block (reference: taskmap.py:57-67, kinematics.py:74-96):
        theta_y = -asin(r20), theta_z = atan2(r10, r00), theta_x = atan2(r21, r22)
    (dividing both atan2 arguments by cos(theta_y) > 0, as the reference does, leaves the angles unchanged).
    Inside a compiled tree the chain [FK, 4x4ToEuler] is the kernels' RMP2_SPACE_FRAME_EULER (analytic
    Euler-rate map).  This stand-alone map of the 16 matrix entries has closed-form derivatives too:
        d theta_y / d r20 = -1 / sqrt(1 - r20^2),   d atan2(y, x) = (x dy - y dx) / (x^2 + y^2),
    and c = qd^T Hess qd with Hess(-asin) = -r / (1 - r^2)^(3/2) and
    Hess(atan2) = [[2xy, y^2 - x^2], [y^2 - x^2, -2xy]] / (x^2 + y^2)^2 in (x, y)."""

    @staticmethod
    def _angles(T):
        r00, r10, r20, r21, r22 = T[:, 0], T[:, 4], T[:, 8], T[:, 9], T[:, 10]
        return torch.stack((torch.atan2(r21, r22), -torch.asin(r20), torch.atan2(r10, r00)), dim=-1)

    def forward(self, input):
        t = to_device(input).reshape(-1, 16)
        return like_input(self._angles(t), input)

    def differentiate(self, q, qd):
        """x [K,3], xd [K,3], J [K,3,16], c [K,3] with q = vec(T) [K,16] and qd its velocity."""
        dev = require_cuda()
        T, Td = to_device(q, dev).reshape(-1, 16), to_device(qd, dev).reshape(-1, 16)
        K = T.shape[0]
        J = torch.zeros(K, 3, 16, device=dev)
        c = torch.zeros(K, 3, device=dev)
        one_m = 1.0 - T[:, 8] * T[:, 8]
        J[:, 1, 8] = -torch.rsqrt(one_m)
        c[:, 1] = -T[:, 8] * Td[:, 8] * Td[:, 8] * one_m.pow(-1.5)
        for row, iy, ix in ((0, 9, 10), (2, 4, 0)):                 # theta_x = atan2(r21, r22), theta_z = atan2(r10, r00)
            x, y, xd_, yd_ = T[:, ix], T[:, iy], Td[:, ix], Td[:, iy]
            rho2 = x * x + y * y
            J[:, row, iy] = x / rho2
            J[:, row, ix] = -y / rho2
            c[:, row] = (2 * x * y * (xd_ * xd_ - yd_ * yd_) + 2 * (y * y - x * x) * xd_ * yd_) / (rho2 * rho2)
        xdot = (J @ Td[..., None])[..., 0]
        return tuple(like_input(o, q) for o in (self._angles(T), xdot, J, c))


class TaskmapRelative4x4(Taskmap):
    """T = T_reference @ Translation(relative_pos_k) for K points fixed in the reference frame
    (reference: taskmap.py:79-99); used by the v1 CollisionAvoidance path
    (experiments/two_joint_robot/05_obstacle_avoidance.py:51-61).  Inside a compiled tree the chain
    [FK, Relative4x4, 4x4ToPosition] is evaluated analytically by the step kernel."""

    def __init__(self, relative_pos):
        self.relative_pos = relative_pos              # [K,3] tensor / array / Variable

    def _rel(self, dev):
        return to_device(self.relative_pos, dev).reshape(-1, 3)

    def forward(self, input):
        dev = require_cuda()
        rel = self._rel(dev)
        T = to_device(input, dev).reshape(-1, 4, 4).expand(rel.shape[0], 4, 4).clone()
        T[:, :3, 3] = T[:, :3, 3] + (T[:, :3, :3] @ rel[:, :, None])[:, :, 0]
        return like_input(T.reshape(-1, 16), input)

    def differentiate(self, q, qd):
        """x [K,16], xd, J [K,16,16], c = 0 (the map is linear in the 16 entries of T_reference)."""
        dev = require_cuda()
        rel = self._rel(dev)
        K = rel.shape[0]
        J = torch.zeros(K, 16, 16, device=dev)
        idx = torch.arange(16, device=dev)
        J[:, idx, idx] = 1.0
        for r in range(3):                          # T[r,3] += sum_c T[r,c] rel[c]
            for c in range(3):
                J[:, 4 * r + 3, 4 * r + c] = rel[:, c]
        x = (J @ to_device(q, dev).reshape(1, 16, 1).expand(K, 16, 1))[:, :, 0]
        xd = (J @ to_device(qd, dev).reshape(1, 16, 1).expand(K, 16, 1))[:, :, 0]
        out = (x, xd, J, torch.zeros(K, 16, device=dev))
        return tuple(like_input(t, q) for t in out)

    def current_points(self):
        return torch.as_tensor(unwrap(self.relative_pos), dtype=torch.float32).reshape(-1, 3).cpu()


def _chain_taskmaps(taskmap_1, taskmap_2):
    """Chain rule for (x, xd, J, c) (reference: taskmap.py:142-162)."""
    def combined_forward(q):
        return taskmap_2.forward(taskmap_1.forward(q))

    def combined_differentiate(q, qd):
        out_1, dout1_dt, J_1, c_1 = taskmap_1.differentiate(q, qd)
        out_2, _, J_2, c_2 = taskmap_2.differentiate(out_1, dout1_dt)
        dev = require_cuda()
        J_1d, J_2d = to_device(J_1, dev), to_device(J_2, dev)
        dout_dt = (J_2d @ to_device(dout1_dt, dev)[..., None])[..., 0]
        J = J_2d @ J_1d
        c = to_device(c_2, dev) + (J_2d @ to_device(c_1, dev)[..., None])[..., 0]
        return tuple(like_input(t, q) for t in (to_device(out_2, dev), dout_dt, J, c))

    stages_1 = taskmap_1.stages if isinstance(taskmap_1, TaskmapByFunction) else [taskmap_1]
    stages_2 = taskmap_2.stages if isinstance(taskmap_2, TaskmapByFunction) else [taskmap_2]
    stages = None if (stages_1 is None or stages_2 is None) else list(stages_1) + list(stages_2)
    return TaskmapByFunction(combined_forward, combined_differentiate, stages=stages)


def chain_taskmaps(taskmap_list):
    """Left fold over the list (reference: taskmap.py:164-168)."""
    chained = taskmap_list[0]
    for taskmap in taskmap_list[1:]:
        chained = _chain_taskmaps(chained, taskmap)
    return chained
