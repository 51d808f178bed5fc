"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A torch-CPU restatement of the reference's control-step hot path
(TomGoesGitHub/Riemannian-Motion-Policies), one function per reference function, each
citing the reference file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline legs may import this module; the product package
(``riemannian_motion_policies_b200``) never does and fails loudly without its CUDA library.

Why a restatement: the reference is TensorFlow 2.10 + PyBullet; neither is installed in the
build container or on the GPU box and there is no network.  The
reference's derivatives come from ``tf.GradientTape``; here they come from ``torch.func``
autodiff (``jvp`` / ``jacrev``), i.e. they are *independent* of the analytic Jacobians the
CUDA kernels use.

PINNING STATUS
  * URDF frame order / constants: pinned against the reference's own, importable
    ``helper/urdf_parsing.py`` (tests/golden/make_golden.py ran it in the build container;
    tests/golden/urdf_frames.json).
  * Rotation helpers, Rodrigues formula, homogeneous transform, Euler extraction: pinned
    against SciPy exactly as the reference's tests do (reference: tests/test_kinematic_forwards.py:16-106).
  * FK of every Panda frame: the reference pins it against PyBullet (tests/test_kinematic_forwards.py:108-137);
    PyBullet is absent, so it is pinned against an independent float64 SciPy FK instead.
  * FK derivatives (x, xd, J, J-dot q-dot), every leaf policy, the task-map chains, the pullback, the
    accumulation and the resolve: pinned against the REFERENCE'S OWN SOURCE FILES, executed unchanged in the
    build container with a TensorFlow-API stand-in over torch (oracle/tf_shim; TensorFlow itself is not
    installed): tests/golden/run_reference_under_shim.py -> tests/golden/ref_*.npz, checked by
    tests/test_reference_golden.py (and re-generated live there whenever /root/reference is present).
  * Still restated rather than executed, because they live inside TensorFlow 2.10 (third party, absent):
    ``tf.linalg.pinv`` (rcond = 10 * max(rows, cols) * eps, SVD, singular values <= rcond * max replaced by
    inf), the ``ndarray += Tensor`` deferral that makes the accumulators float32 and
    ``GradientTape`` (torch autograd in the shim).  For exactly these three, parity with a real TensorFlow
    run remains UNPINNED; the reference holds no test or golden vector for them.

dtype switch: ``dtype=torch.float32`` is the reference-faithful mode (everything, including the
accumulators and the pinv, runs in float32 -- SURVEY.md section 0); ``dtype=torch.float64`` is the
"truth" mode: float64 arithmetic with the *float32* pinv cutoff (the cutoff is part of the
semantics, not of the rounding).
"""
from xml.etree import ElementTree

import numpy as np
import torch
from torch.func import jacrev, jvp

EPS32 = float(np.finfo(np.float32).eps)


# =============================================================================================
# helper/urdf_parsing.py
# =============================================================================================
def urdf_frames(filepath):
    """Frames of a URDF in the reference's order.

    reference: helper/urdf_parsing.py:57-97 (tree built breadth first from the one link that is no
    joint's child, every pass scanning all joints in file order) and :134-147 (one backward path
    per non-root element, in id order) -- ids are handed out in creation order, so the frame
    order is the creation order.  Returns a list of dicts with the fields of ``UrdfElem``
    (helper/urdf_parsing.py:4-16) plus ``path`` = joint names from the base to the frame.
    """
    root = ElementTree.parse(filepath).getroot()
    all_links = root.findall("link")
    all_joints = root.findall("joint")
    base = None
    for link in all_links:
        if not any(j.find("child").attrib["link"] == link.attrib["name"] for j in all_joints):
            base = link.attrib["name"]
            break
    elems = [dict(name="<ROOT>", link_name=base, path=[])]
    todo = [0]
    while todo:
        leaf = elems[todo.pop(0)]
        for joint in all_joints:
            if joint.find("parent").attrib["link"] != leaf["link_name"]:
                continue
            child_name = joint.find("child").attrib["link"]
            child_link = next(ln for ln in all_links if ln.attrib["name"] == child_name)
            jtype = joint.attrib["type"]
            col = child_link.find("collision")
            elems.append(dict(
                name=joint.attrib["name"], link_name=child_name, joint_type=jtype,
                rpy=[float(v) for v in joint.find("origin").attrib["rpy"].split()],
                xyz=[float(v) for v in joint.find("origin").attrib["xyz"].split()],
                axis=([float(v) for v in joint.find("axis").attrib["xyz"].split()]
                      if jtype != "fixed" else [0., 0., 0.]),
                # ``True if link.find('collision') else False`` == element exists and has children
                has_collision=bool(col is not None and len(col) > 0),
                path=leaf["path"] + [joint.attrib["name"]]))
            todo.append(len(elems) - 1)
    return elems[1:]


# =============================================================================================
# kinematics.py
# =============================================================================================
def reduce_matrix_prod(all_T):
    """Left-to-right product of a stack of 4x4 matrices.  reference: kinematics.py:12-20."""
    m = torch.eye(4, dtype=all_T.dtype)
    for i in range(all_T.shape[0]):
        m = m @ all_T[i]
    return m


def R_x(angle):
    """angle [B,1] -> [B,3,3].  reference: kinematics.py:22-32."""
    c, s = torch.cos(angle), torch.sin(angle)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([torch.cat([o, z, z], -1), torch.cat([z, c, -s], -1), torch.cat([z, s, c], -1)], dim=-2)


def R_y(angle):
    """reference: kinematics.py:34-44."""
    c, s = torch.cos(angle), torch.sin(angle)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([torch.cat([c, z, s], -1), torch.cat([z, o, z], -1), torch.cat([-s, z, c], -1)], dim=-2)


def R_z(angle):
    """reference: kinematics.py:46-56."""
    c, s = torch.cos(angle), torch.sin(angle)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([torch.cat([c, -s, z], -1), torch.cat([s, c, z], -1), torch.cat([z, z, o], -1)], dim=-2)


def homogenous_transformation(R, t):
    """R [B,3,3], t [B,3] -> [B,4,4].  reference: kinematics.py:58-71."""
    assert R.shape[-2:] == (3, 3) and t.shape[-1] == 3
    Rt = torch.cat([R, t[..., None]], dim=-1)
    bottom = torch.cat([torch.zeros(R.shape[0], 1, 3, dtype=R.dtype), torch.ones(R.shape[0], 1, 1, dtype=R.dtype)], -1)
    return torch.cat([Rt, bottom], dim=-2)


def euler_from_rotation_matrix(rotation_matrix):
    """xyz Euler angles.  reference: kinematics.py:74-96."""
    r00, r10 = rotation_matrix[:, 0, 0], rotation_matrix[:, 1, 0]
    r21, r22, r20 = rotation_matrix[:, 2, 1], rotation_matrix[:, 2, 2], rotation_matrix[:, 2, 0]
    theta_y = -torch.asin(r20)
    cos_y = torch.cos(theta_y)
    safe = torch.where(torch.abs(cos_y) < 1e-6, torch.ones_like(cos_y), cos_y)
    theta_z = torch.atan2(r10 / safe, r00 / safe)
    theta_x = torch.atan2(r21 / safe, r22 / safe)
    return torch.stack((theta_x, theta_y, theta_z), dim=-1)


def rotation_matrix_from_rotation_vector(vec, angle):
    """Rodrigues with the axis used as given (not normalised).  reference: kinematics.py:99-121."""
    assert vec.shape[0] == angle.shape[0]
    cos = torch.cos(angle)[:, None, None]
    sin = torch.sin(angle)[:, None, None]
    vec0 = torch.cat([torch.zeros(vec.shape[0], 1, dtype=vec.dtype), vec], dim=-1)
    eye = torch.eye(3, dtype=vec.dtype).expand(vec.shape[0], 3, 3)
    outer = vec[:, :, None] * vec[:, None, :]
    sign = torch.tensor([[1, -1, 1], [1, 1, -1], [-1, 1, 1]], dtype=vec.dtype)
    where = torch.tensor([[0, 3, 2], [3, 0, 1], [2, 1, 0]])
    u_tilde = sign * vec0[:, where]
    return cos * eye + sin * u_tilde + (1 - cos) * outer


def rotation_matrix_from_rpy(rpy):
    """``R_x(roll) @ R_y(pitch) @ R_z(yaw)``.  reference: kinematics.py:123-127."""
    roll, pitch, yaw = rpy[:, 0:1], rpy[:, 1:2], rpy[:, 2:3]
    return R_x(roll) @ R_y(pitch) @ R_z(yaw)


class UrdfForwardKinematic:
    """reference: kinematics.py:155-270."""

    def __init__(self, urdf_filepath, order, dtype=torch.float32):
        self.filepath, self.order, self.n_joints, self.dtype = urdf_filepath, list(order), len(order), dtype
        self._build()

    def _build(self):
        """reference: kinematics.py:163-210.  Constants are built in float32 (tf.constant of
        Python floats) and only then widened when dtype is float64, so both modes share them."""
        frames = urdf_frames(self.filepath)
        keys = [f["name"] for f in frames]
        self.frame_names = keys
        self._name_to_idx = {k: i for i, k in enumerate(keys)}
        F = len(keys)
        max_len = max(len(f["path"]) for f in frames)
        self.kinematic_chains = torch.tensor(
            [[self._name_to_idx[p] for p in f["path"]] + [F] * (max_len - len(f["path"])) for f in frames])
        self._q_reordering = torch.tensor([self.order.index(k) if k in self.order else len(self.order) for k in keys])
        rpy = torch.tensor([f["rpy"] for f in frames], dtype=torch.float32)
        xyz = torch.tensor([f["xyz"] for f in frames], dtype=torch.float32)
        self.T_constant = homogenous_transformation(rotation_matrix_from_rpy(rpy), xyz).to(self.dtype)
        self.axis = torch.tensor([f["axis"] for f in frames], dtype=torch.float32).to(self.dtype)
        jt = [f["joint_type"] for f in frames]
        as_mask = lambda kind: torch.tensor([1.0 if t == kind else 0.0 for t in jt], dtype=self.dtype)[:, None, None]
        self.is_revolute, self.is_prismatic, self.is_fixed = as_mask("revolute"), as_mask("prismatic"), as_mask("fixed")
        self.has_collision = [f["has_collision"] for f in frames]

    def forward(self, q, frame):
        """q [1,n] -> T [1,4,4].  reference: kinematics.py:212-247."""
        q = q.reshape(-1).to(self.dtype)
        F = self.T_constant.shape[0]
        q = torch.cat([q, torch.zeros(1, dtype=self.dtype)])
        q = q[self._q_reordering]
        T_fixed = torch.eye(4, dtype=self.dtype).expand(F, 4, 4)
        R_rev = rotation_matrix_from_rotation_vector(self.axis.reshape(-1, 3), q.reshape(-1))
        T_rev = homogenous_transformation(R_rev, torch.zeros(F, 3, dtype=self.dtype))
        R_pri = torch.eye(3, dtype=self.dtype).expand(F, 3, 3)
        T_pri = homogenous_transformation(R_pri, q[:, None] * self.axis)
        T_var = self.is_fixed * T_fixed + self.is_revolute * T_rev + self.is_prismatic * T_pri
        T = self.T_constant @ T_var
        T = torch.cat([T, torch.eye(4, dtype=self.dtype)[None]], dim=0)
        chain = self.kinematic_chains[self._name_to_idx[frame]]
        return reduce_matrix_prod(T[chain])[None]

    __call__ = forward   # taskmap.py:28 calls ``self.fkine(q, self.frame)``

    def differentiate(self, q, qd, frame):
        """-> x [1,16], xd [1,16], J [1,16,n], c [1,16].  reference: kinematics.py:250-270
        (xd = J qd by a Jacobian-vector product, J by reverse mode, c = d(J qd)/dq . qd)."""
        q = q.reshape(-1).to(self.dtype)
        qd = qd.reshape(-1).to(self.dtype)
        fn = lambda q_: self.forward(q_[None], frame).reshape(-1)
        vel = lambda q_: jvp(fn, (q_,), (qd,))[1]
        x, xd = jvp(fn, (q,), (qd,))
        J = jacrev(fn)(q)
        c = jvp(vel, (q,), (qd,))[1]
        return x[None], xd[None], J[None], c[None]


# =============================================================================================
# helper/rmp_helper.py
# =============================================================================================
def rmp_differentiate(fn):
    """(x, xd, J, c) of a row-wise forward function.  reference: helper/rmp_helper.py:3-22.
    ``batch_jacobian`` assumes independent rows, so J[k] = d(sum_k' x[k'])/dq[k]."""
    def differentiate_fn(q, qd):
        vel = lambda q_: jvp(fn, (q_,), (qd,))[1]
        x, xd = jvp(fn, (q,), (qd,))
        c = jvp(vel, (q,), (qd,))[1]
        J = jacrev(lambda q_: fn(q_).sum(0))(q).permute(1, 0, 2)      # [m,K,p] -> [K,m,p]
        return x, xd, J, c
    return differentiate_fn


def soft_norm(v, c):
    """reference: helper/rmp_helper.py:62-65."""
    n = torch.linalg.norm(v, dim=-1)
    h = n + 1 / c * torch.log(1 + torch.exp(-2 * c * n))
    return v / h[:, None]


def directionally_stretched_metric(v, beta, c):
    """reference: helper/rmp_helper.py:67-74."""
    zeta = soft_norm(v, c)
    A = zeta[..., :, None] * zeta[..., None, :]
    eye = torch.eye(A.shape[-1], dtype=v.dtype).expand(A.shape[0], -1, -1)
    return beta * A + (1 - beta) * eye


# =============================================================================================
# taskmap.py
# =============================================================================================
class IdentityTaskmap:
    """reference: taskmap.py:13-20."""
    def forward(self, q):
        return q

    def differentiate(self, q, qd):
        return rmp_differentiate(self.forward)(q, qd)


class TaskmapByForwardKinematic:
    """reference: taskmap.py:22-31."""
    def __init__(self, fkine, frame):
        self.fkine, self.frame = fkine, frame

    def forward(self, q):
        return self.fkine(q, self.frame)

    def differentiate(self, q, qd):
        return self.fkine.differentiate(q, qd, self.frame)


class TaskmapByFunction:
    """reference: taskmap.py:33-42."""
    def __init__(self, forward_fn, differentiate_fn):
        self.forward_fn, self.differentiate_fn = forward_fn, differentiate_fn

    def forward(self, q):
        return self.forward_fn(q)

    def differentiate(self, q, qd):
        return self.differentiate_fn(q, qd)


class TaskmapFrom4x4ToPosition:
    """reference: taskmap.py:45-54."""
    def forward(self, input):
        return input.reshape(-1, 4, 4)[:, :3, 3]

    def differentiate(self, q, qd):
        return rmp_differentiate(self.forward)(q, qd)


class TaskmapFrom4x4ToEuler:
    """reference: taskmap.py:57-67."""
    def forward(self, input):
        return euler_from_rotation_matrix(input.reshape(-1, 4, 4)[:, :3, :3])

    def differentiate(self, q, qd):
        return rmp_differentiate(self.forward)(q, qd)


class TaskmapRelative4x4:
    """reference: taskmap.py:79-99."""
    def __init__(self, relative_pos):
        self.relative_pos = relative_pos

    def forward(self, input):
        K = self.relative_pos.shape[0]
        T_ref = input.reshape(-1, 4, 4).expand(K, 4, 4)
        T_rel = homogenous_transformation(torch.eye(3, dtype=input.dtype).expand(K, 3, 3), self.relative_pos.to(input.dtype))
        return (T_ref @ T_rel).reshape(-1, 16)

    def differentiate(self, q, qd):
        K = self.relative_pos.shape[0]
        return rmp_differentiate(self.forward)(q.repeat_interleave(K, 0), qd.repeat_interleave(K, 0))


class TaskmapJointFrame4x4ToDistance:
    """reference: taskmap.py:115-138 (gradient flows only through the frame origin)."""
    def __init__(self, pos_on_link_in_base_frame, pos_on_obstacle_in_base_frame):
        self.pos_on_link_in_base_frame = pos_on_link_in_base_frame
        self.pos_on_obstacle_in_base_frame = pos_on_obstacle_in_base_frame

    def forward(self, input):
        link = self.pos_on_link_in_base_frame.to(input.dtype)
        obst = self.pos_on_obstacle_in_base_frame.to(input.dtype)
        T_ref = input.reshape(-1, 4, 4).expand(link.shape[0], 4, 4)
        pos_joint = T_ref[:, :3, 3]
        rel = (link - pos_joint).detach()                      # tf.stop_gradient (taskmap.py:126)
        critical = pos_joint + rel
        return torch.linalg.norm(critical - obst, dim=-1)[:, None]

    def differentiate(self, q, qd):
        K = self.pos_on_link_in_base_frame.shape[0]
        return rmp_differentiate(self.forward)(q.repeat_interleave(K, 0), qd.repeat_interleave(K, 0))


def _chain_taskmaps(taskmap_1, taskmap_2):
    """reference: taskmap.py:142-162."""
    def combined_forward(q):
        return taskmap_2.forward(taskmap_1.forward(q))

    def combined_differentiate(q, qd):
        out_1, dout1_dt, J_1, c_1 = taskmap_1.differentiate(q, qd)
        out_2, _, J_2, c_2 = taskmap_2.differentiate(out_1, dout1_dt)
        dout_dt = (J_2 @ dout1_dt[..., None])[..., 0]
        J = J_2 @ J_1
        c = c_2 + (J_2 @ c_1[..., None])[..., 0]
        return out_2, dout_dt, J, c

    return TaskmapByFunction(combined_forward, combined_differentiate)


def chain_taskmaps(taskmap_list):
    """reference: taskmap.py:164-168."""
    chained = taskmap_list[0]
    for tm in taskmap_list[1:]:
        chained = _chain_taskmaps(chained, tm)
    return chained


# =============================================================================================
# third party: tf.linalg.pinv (TensorFlow 2.10.0, python/ops/linalg/linalg_impl.py)
# =============================================================================================
def tf_pinv(a, eps=EPS32):
    """rcond = 10 * max(rows, cols) * eps; singular values <= rcond * max(s) become inf;
    pinv = (V / s) @ U^H.  Call site: rmp.py:153."""
    rcond = 10.0 * max(a.shape[-2], a.shape[-1]) * eps
    u, s, vh = torch.linalg.svd(a, full_matrices=False)
    cutoff = rcond * s.max(dim=-1).values
    s = torch.where(s > cutoff[..., None], s, torch.full_like(s, float("inf")))
    return (vh.transpose(-1, -2) / s[..., None, :]) @ u.transpose(-1, -2)


# =============================================================================================
# rmp.py
# =============================================================================================
class RmpCore:
    """reference: rmp.py:111-180."""

    def __init__(self, rmps=None, dtype=torch.float32):
        self.rmps = {} if rmps is None else rmps
        self.dtype = dtype

    def add_rmp(self, rmp):
        self.rmps[rmp.name] = rmp

    def remove_rmp_by_name(self, name):
        self.rmps.pop(name)

    def combine(self, q, qd):
        """Sum of pulled-back leaves (f, M) before the resolve.  reference: rmp.py:135-150.
        The accumulators take the tensors' dtype after the first ``+=``."""
        q = torch.as_tensor(q).to(self.dtype).reshape(-1)
        qd = torch.as_tensor(qd).to(self.dtype).reshape(-1)
        n = q.shape[0]
        f_combined = torch.zeros(n, dtype=self.dtype)
        M_combined = torch.zeros(n, n, dtype=self.dtype)
        for rmp in self.rmps.values():
            f, M = self._calculate_rmp(rmp, q, qd)
            f_combined = f_combined + f.sum(0)
            M_combined = M_combined + M.sum(0)
        return f_combined, M_combined

    def evaluate(self, q, qd):
        """reference: rmp.py:133-155."""
        f_combined, M_combined = self.combine(q, qd)
        return tf_pinv(M_combined) @ f_combined

    def _calculate_rmp(self, rmp, q, qd):
        """reference: rmp.py:157-180."""
        x, xd, J, c = rmp.taskmap.differentiate(q[None, :], qd[None, :])
        xdd_des, M_leaf = rmp.evaluate(x, xd)
        Jt = J.transpose(1, 2)
        f = ((Jt @ M_leaf) @ (xdd_des - c)[..., None])[..., 0]
        M = (Jt @ M_leaf) @ J
        return f, M


class RiemannianMotionPolicy:
    """reference: rmp.py:184-206 / rmp2.py:6-29."""
    def __init__(self, name, taskmap):
        self.name, self.taskmap = name, taskmap

    def evaluate(self, x, xd):
        return self._motion_command(x, xd), self._metric(x, xd)


class TargetPolicy(RiemannianMotionPolicy):
    """reference: rmp.py:226-261."""
    def __init__(self, alpha, beta, c, goal, taskmap, name="Target_RMP"):
        super().__init__(name, taskmap)
        self.goal, self.c, self.alpha, self.beta = goal, c, alpha, beta
        self.sigma_H, self.sigma_w = 1, 3

    def _goal(self, x):
        return torch.as_tensor(self.goal).to(x.dtype)

    def _motion_command(self, x, xd):
        v = self._goal(x) - x
        nv = torch.linalg.norm(v)                                   # whole-tensor norm (rmp.py:243)
        h = nv + self.c * torch.log(1 + torch.exp(-2 * self.c * nv))
        return self.alpha * (1 / h * v) - self.beta * xd

    def _metric(self, x, xd):
        f_attract = self._motion_command(x, xd)
        nv = torch.linalg.norm(x - self._goal(x))
        beta = 1 - torch.exp(-0.5 * nv ** 2 / self.sigma_H ** 2)
        H = directionally_stretched_metric(v=f_attract, c=self.c, beta=beta)
        w = torch.exp(-nv / self.sigma_w)
        return w * H


class CollisionAvoidance(RiemannianMotionPolicy):
    """reference: rmp.py:264-315 (v1 obstacle leaf; d [K] and vec [K,3] are external data)."""
    def __init__(self, d, vec, eta_rep, nu_rep, eta_damp, nu_damp, r, c, taskmap, name="collision_avoidance"):
        super().__init__(name, taskmap)
        self.d, self.vec = d, vec
        self.eta_rep, self.nu_rep, self.eta_damp, self.nu_damp, self.r, self.c = eta_rep, nu_rep, eta_damp, nu_damp, r, c

    def _motion_command(self, x, xd):
        d, vec = torch.as_tensor(self.d).to(x.dtype), torch.as_tensor(self.vec).to(x.dtype)
        alpha_rep = self.eta_rep * torch.exp(-d / self.nu_rep)
        f_rep = alpha_rep[:, None] * vec
        epsilon = 1e-6
        alpha_damp = self.eta_damp / (d / self.nu_damp + epsilon)
        scaling = torch.clamp((-xd * vec).sum(-1), min=0.)
        P_obs = scaling[:, None, None] * vec[:, :, None] * vec[:, None, :]
        f_damp = alpha_damp[:, None] * (P_obs @ xd[..., None])[..., 0]
        return f_rep - f_damp

    def _metric(self, x, xd):
        d = torch.as_tensor(self.d).to(x.dtype)
        c_0, c_1, c_2, c_3 = 1, 0, -3 / self.r ** 2, 2 / self.r ** 3
        spline = c_3 * d ** 3 + c_2 * d ** 2 + c_1 * d + c_0
        w = torch.where(d > self.r, torch.zeros_like(spline), spline)
        f_obs = self._motion_command(x, xd)
        H = directionally_stretched_metric(v=f_obs, c=self.c, beta=0)
        return w[:, None, None] * H


class ConfigurationSpaceBiasing(RiemannianMotionPolicy):
    """reference: rmp.py:318-347."""
    def __init__(self, gamma_p, gamma_d, q0, name, w=0.05):
        super().__init__(name, IdentityTaskmap())
        self.gamma_p, self.gamma_d, self.q_0, self.w = gamma_p, gamma_d, q0, w

    def _motion_command(self, x, xd):
        return self.gamma_p * (torch.as_tensor(self.q_0).to(x.dtype) - x) - self.gamma_d * xd

    def _metric(self, x, xd):
        return self.w * torch.eye(x.shape[-1], dtype=x.dtype)[None]


class JointLimitAvoidance(RiemannianMotionPolicy):
    """reference: rmp.py:349-382.  ``w * H`` broadcasts w over the last axis (column scaling)."""
    def __init__(self, lower_limits, upper_limits, gamma_p, gamma_d, name="joint_limit_avoidance"):
        super().__init__(name, IdentityTaskmap())
        self.lower_limits = torch.as_tensor(np.asarray(lower_limits), dtype=torch.float32)
        self.upper_limits = torch.as_tensor(np.asarray(upper_limits), dtype=torch.float32)
        self.gamma_p, self.gamma_d = gamma_p, gamma_d

    def _metric(self, q, qd):
        lo, up = self.lower_limits.to(q.dtype), self.upper_limits.to(q.dtype)
        d_upper = (up - q) / (up - lo)
        d_lower = (q - lo) / (up - lo)
        d = torch.minimum(d_upper, d_lower)
        r = 0.15
        c_0, c_1, c_2, c_3 = 1, 0, -3 / r ** 2, 2 / r ** 3
        spline = c_3 * d ** 3 + c_2 * d ** 2 + c_1 * d + c_0
        w = torch.where(d > r, torch.zeros_like(spline), spline)
        qd_max = 20 * (2 * np.pi) / 60
        H = directionally_stretched_metric(qd / qd_max, beta=0.9, c=5)
        return w * H

    def _motion_command(self, q, qd):
        return -self.gamma_p * q - self.gamma_d * qd


# =============================================================================================
# rmp2.py
# =============================================================================================
class TargetAttractor(RiemannianMotionPolicy):
    """reference: rmp2.py:31-83."""
    def __init__(self, goal, accel_p_gain, accel_d_gain, accel_norm_eps, metric_alpha_length_scale,
                 min_metric_alpha, max_metric_scalar, min_metric_scalar, proximity_metric_boost_scalar,
                 proximity_metric_boost_length_scale, taskmap, name="attractor"):
        super().__init__(name, taskmap)
        self.goal = goal
        self.accel_p_gain, self.accel_d_gain, self.accel_norm_eps = accel_p_gain, accel_d_gain, accel_norm_eps
        self.metric_alpha_length_scale, self.min_metric_alpha = metric_alpha_length_scale, min_metric_alpha
        self.max_metric_scalar, self.min_metric_scalar = max_metric_scalar, min_metric_scalar
        self.proximity_metric_boost_scalar = proximity_metric_boost_scalar
        self.proximity_metric_boost_length_scale = proximity_metric_boost_length_scale

    def _motion_command(self, x, xd):
        delta = torch.as_tensor(self.goal).to(x.dtype) - x
        delta_norm = torch.linalg.norm(delta, dim=1)[:, None]
        return self.accel_p_gain * delta / (delta_norm + self.accel_norm_eps) - self.accel_d_gain * xd

    def _metric(self, x, xd):
        delta = torch.as_tensor(self.goal).to(x.dtype) - x
        delta_norm = torch.linalg.norm(delta, dim=1)[:, None]
        soft = torch.maximum(delta_norm, self.accel_norm_eps / 10 * torch.ones_like(delta_norm))
        delta_hat = delta / soft
        eye = torch.eye(x.shape[1], dtype=x.dtype)[None]
        S = delta_hat[:, :, None] * delta_hat[:, None, :]
        scaled = delta_norm / self.metric_alpha_length_scale
        a = ((1. - self.min_metric_alpha) * torch.exp(-.5 * scaled * scaled) + self.min_metric_alpha)[..., None]
        metric = a * self.max_metric_scalar * eye + (1. - a) * self.min_metric_scalar * S
        bscaled = delta_norm / self.proximity_metric_boost_length_scale
        boost_a = torch.exp(-.5 * bscaled * bscaled)
        boost = (boost_a * self.proximity_metric_boost_scalar + (1. - boost_a) * 1.)[..., None]
        return boost * metric


class JointVelocityCap(RiemannianMotionPolicy):
    """reference: rmp2.py:86-112 (the ``tf.where`` on line 107 is discarded; the metric divides
    the dense matrix element-wise)."""
    def __init__(self, max_velocity, velocity_damping_region, damping_gain, metric_weight, name="joint_velocity_cap"):
        super().__init__(name, IdentityTaskmap())
        self.max_velocity, self.velocity_damping_region = max_velocity, velocity_damping_region
        self.damping_gain, self.metric_weight, self.eps = damping_gain, metric_weight, 1e-6
        self.damped_velocity_cutoff = max_velocity - velocity_damping_region

    def evaluate(self, x, xd):
        delta_velocity = torch.abs(xd) - self.damped_velocity_cutoff
        xdd = -torch.abs(self.damping_gain * delta_velocity) * torch.sign(xd)
        clipped = torch.clamp(delta_velocity, max=self.velocity_damping_region - self.eps)
        ratio = clipped / self.velocity_damping_region
        diag = torch.diag_embed(ratio ** 2)
        metric = self.metric_weight / (1.0 - diag)
        accel = torch.where(torch.abs(xd) < self.damped_velocity_cutoff, torch.zeros_like(xdd), xdd)
        return accel, metric


class JointDamping(RiemannianMotionPolicy):
    """reference: rmp2.py:115-137."""
    def __init__(self, accel_d_gain, metric_scalar, inertia, name="joint_damping"):
        super().__init__(name, IdentityTaskmap())
        self.accel_d_gain, self.metric_scalar, self.inertia = accel_d_gain, metric_scalar, inertia

    def evaluate(self, x, xd):
        xd_norm = torch.linalg.norm(xd, dim=1, keepdim=True)
        accel = -(self.accel_d_gain * xd_norm) * xd
        scal = (self.metric_scalar * xd_norm)[..., None]
        metric = torch.eye(x.shape[1], dtype=x.dtype)[None] * (scal + self.inertia)
        return accel, metric


class ObstacleAvoidance(RiemannianMotionPolicy):
    """reference: rmp2.py:140-196."""
    def __init__(self, margin, damping_gain, damping_std_dev, damping_robustness_eps,
                 damping_velocity_gate_length_scale, repulsion_gain, repulsion_std_dev,
                 metric_modulation_radius, metric_scalar, metric_exploder_std_dev, metric_exploder_eps,
                 taskmap, name):
        super().__init__(name, taskmap)
        self.margin, self.damping_gain, self.damping_std_dev = margin, damping_gain, damping_std_dev
        self.damping_robustness_eps = damping_robustness_eps
        self.damping_velocity_gate_length_scale = damping_velocity_gate_length_scale
        self.repulsion_gain, self.repulsion_std_dev = repulsion_gain, repulsion_std_dev
        self.metric_modulation_radius, self.metric_scalar = metric_modulation_radius, metric_scalar
        self.metric_exploder_std_dev, self.metric_exploder_eps = metric_exploder_std_dev, metric_exploder_eps

    def evaluate(self, x, xd):
        r = self.metric_modulation_radius
        x = x - self.margin
        x = torch.maximum(x, torch.zeros_like(x))
        base_metric = self.metric_scalar / (x / self.metric_exploder_std_dev + self.metric_exploder_eps)
        gate = x * x / (r * r) - 2. * x / r + 1.
        gate = torch.where(x > r, torch.zeros_like(gate), gate)
        metric = base_metric * gate
        xdd_repel = self.repulsion_gain * torch.exp(-(x / self.repulsion_std_dev))
        sig = torch.sigmoid(xd / self.damping_velocity_gate_length_scale)
        xdd_damping = -(1. - sig) * self.damping_gain * xd / (x / self.damping_std_dev + self.damping_robustness_eps)
        accel = xdd_repel + xdd_damping
        metric = torch.where(x > r, torch.zeros_like(metric), (1 - sig) * metric)
        return accel, metric[..., None]


class CSpaceBiasing(RiemannianMotionPolicy):
    """reference: rmp2.py:198-226."""
    def __init__(self, goal, metric_scalar, position_gain, damping_gain, robust_position_term_thresh,
                 inertia, taskmap=None, name="cspace_target"):
        super().__init__(name, IdentityTaskmap() if taskmap is None else taskmap)
        self.goal, self.metric_scalar, self.position_gain = goal, metric_scalar, position_gain
        self.damping_gain, self.robust_position_term_thresh, self.inertia = damping_gain, robust_position_term_thresh, inertia

    def evaluate(self, x, xd):
        x = x - torch.as_tensor(self.goal).to(x.dtype)
        x_norm = torch.linalg.norm(x, dim=1, keepdim=True)
        x_hat = x / x_norm
        thr = self.robust_position_term_thresh
        qdd_position = torch.where(x_norm < thr, -x * self.position_gain, -thr * x_hat * self.position_gain)
        metric = torch.eye(x.shape[1], dtype=x.dtype)[None] * (self.metric_scalar + self.inertia)
        return qdd_position + (-self.damping_gain * xd), metric
