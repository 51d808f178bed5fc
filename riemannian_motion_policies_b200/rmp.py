"""RMP root and v1 leaf policies -- host-side mirror of the reference's ``rmp.py``.

``RmpCore.evaluate(q, qd)`` keeps the reference's signature (rmp.py:133) and additionally accepts
``q`` / ``qd`` of shape [B, n] together with per-environment goals and sphere obstacles.  It
compiles the leaf list into kernel tables once (``rmp2_tree_create``) and then runs the step's kernels
(``rmp2_step``: frames -> spheres -> step [-> resolve] -> resolve fallback).  Unsupported (task map, leaf) combinations raise
``NotImplementedError`` -- there is no CPU or generic-autodiff fallback.
"""
import ctypes

import numpy as np
import torch

from . import _native
from ._leaf import RiemannianMotionPolicy, as_float_list
from ._tensor import current_stream_ptr, is_device_tensor, require_cuda, stage_host_inputs, to_device, unwrap
from .taskmap import (IdentityTaskmap, TaskmapByForwardKinematic, TaskmapByFunction, TaskmapFrom4x4ToEuler,
                      TaskmapFrom4x4ToPosition, TaskmapJointFrame4x4ToDistance, TaskmapJointFrame4x4ToSphereDistance,
                      TaskmapRelative4x4)


# =================================================================================================
# leaves of rmp.py
# =================================================================================================
class TargetPolicy(RiemannianMotionPolicy):
    """Move a reference frame (or the joints) to a goal (reference: rmp.py:226-261)."""
    leaf_type = _native.LEAF_TARGET_POLICY

    def __init__(self, alpha, beta, c, goal, taskmap, name='Target_RMP'):
        super().__init__(name, taskmap)
        self.goal = goal
        self.c = c
        self.alpha = alpha
        self.beta = beta
        self.sigma_H = 1
        self.sigma_w = 3

    def _params(self):
        return [self.alpha, self.beta, self.c]

    def _vec(self, dim):
        return as_float_list(self.goal, dim, "TargetPolicy.goal")


class CollisionAvoidance(RiemannianMotionPolicy):
    """v1 obstacle avoidance (reference: rmp.py:264-315).  ``d`` [K] and ``vec`` [K,3] are the distance and
    normal of every closest-point pair, typically Datamanager variables that are updated in place."""
    leaf_type = _native.LEAF_COLLISION_AVOIDANCE

    def __init__(self, d, vec, eta_rep, nu_rep, eta_damp, nu_damp, r, c, taskmap, name='collision_avoidance'):
        super().__init__(name, taskmap)
        self.d, self.vec = d, vec
        self.eta_rep, self.nu_rep, self.eta_damp, self.nu_damp = eta_rep, nu_rep, eta_damp, nu_damp
        self.r, self.c = r, c

    def _params(self):
        return [self.eta_rep, self.nu_rep, self.eta_damp, self.nu_damp, self.r, self.c]

    def current_data(self):
        """[K,4] = (distance, normal xyz) read from the holders right now."""
        d = torch.as_tensor(unwrap(self.d), dtype=torch.float32).reshape(-1, 1).cpu()
        vec = torch.as_tensor(unwrap(self.vec), dtype=torch.float32).reshape(-1, 3).cpu()
        if d.shape[0] != vec.shape[0]:
            raise ValueError("CollisionAvoidance: d and vec need the same number of rows")
        return torch.cat([d, vec], dim=1)

    def _aux(self, K, dev):
        aux = self.current_data()
        if aux.shape[0] != K:
            raise ValueError(f"CollisionAvoidance holds {aux.shape[0]} pairs but x has {K} rows")
        return aux.to(dev).contiguous()


class ConfigurationSpaceBiasing(RiemannianMotionPolicy):
    """PD controller towards q0 (reference: rmp.py:318-347)."""
    leaf_type = _native.LEAF_CONFIG_BIASING

    def __init__(self, gamma_p, gamma_d, q0, name, w=0.05):
        super().__init__(name, taskmap=IdentityTaskmap())
        self.gamma_p = gamma_p
        self.gamma_d = gamma_d
        self.q_0 = q0
        self.w = w

    def _params(self):
        return [self.gamma_p, self.gamma_d, self.w]

    def _vec(self, dim):
        return as_float_list(self.q_0, dim, "ConfigurationSpaceBiasing.q0")


class JointLimitAvoidance(RiemannianMotionPolicy):
    """reference: rmp.py:349-382."""
    leaf_type = _native.LEAF_JOINT_LIMIT

    def __init__(self, lower_limits, upper_limits, gamma_p, gamma_d, name='joint_limit_avoidance'):
        super().__init__(name, taskmap=IdentityTaskmap())
        self.lower_limits = torch.as_tensor(np.asarray(unwrap(lower_limits), dtype=np.float64), dtype=torch.float32)
        self.upper_limits = torch.as_tensor(np.asarray(unwrap(upper_limits), dtype=np.float64), dtype=torch.float32)
        self.gamma_p = gamma_p
        self.gamma_d = gamma_d

    def _params(self):
        return [self.gamma_p, self.gamma_d]

    def _vec(self, dim):
        return (as_float_list(self.lower_limits, dim, "JointLimitAvoidance.lower_limits") +
                as_float_list(self.upper_limits, dim, "JointLimitAvoidance.upper_limits"))


# =================================================================================================
# tree compilation
# =================================================================================================
def classify_taskmap(taskmap):
    """-> (space, fkine, frame_name, distance_taskmap) for the chains the kernel implements."""
    if isinstance(taskmap, IdentityTaskmap):
        return _native.SPACE_CONFIG, None, None, None
    stages = taskmap.stages if isinstance(taskmap, TaskmapByFunction) else None
    if stages is not None and len(stages) == 2 and isinstance(stages[0], TaskmapByForwardKinematic):
        fk, second = stages
        if isinstance(second, TaskmapFrom4x4ToPosition):
            return _native.SPACE_FRAME_POSITION, fk.fkine, fk.frame, None
        if isinstance(second, TaskmapFrom4x4ToEuler):
            return _native.SPACE_FRAME_EULER, fk.fkine, fk.frame, None
        if isinstance(second, TaskmapJointFrame4x4ToSphereDistance):
            return _native.SPACE_FRAME_DISTANCE_SPHERES, fk.fkine, fk.frame, second
        if isinstance(second, TaskmapJointFrame4x4ToDistance):
            return _native.SPACE_FRAME_DISTANCE_PAIRS, fk.fkine, fk.frame, second
    if (stages is not None and len(stages) == 3 and isinstance(stages[0], TaskmapByForwardKinematic)
            and isinstance(stages[1], TaskmapRelative4x4) and isinstance(stages[2], TaskmapFrom4x4ToPosition)):
        return _native.SPACE_FRAME_POINTS, stages[0].fkine, stages[0].frame, stages[1]
    raise NotImplementedError(
        f"task map {type(taskmap).__name__} (stages={[type(s).__name__ for s in stages] if stages else None}) is not "
        "one of the chains the CUDA engine implements: IdentityTaskmap, [FK, 4x4ToPosition], [FK, 4x4ToEuler], "
        "[FK, JointFrame4x4ToDistance], "
        "[FK, JointFrame4x4ToSphereDistance], [FK, Relative4x4, 4x4ToPosition]")


class CompiledTree:
    """Native handle (``rmp2_tree``) of one leaf list, plus what is needed to refresh parameters."""

    def __init__(self, rmps, n, goal_leaves):
        self.n = n
        self.names = list(rmps.keys())
        self.goal_leaves = list(goal_leaves)                      # leaf names fed per environment
        self.fkine = None
        self.entries = []                                         # (leaf, space, frame_idx, goal_slot, dist_taskmap)
        for name, leaf in rmps.items():
            space, fkine, frame, dist = classify_taskmap(leaf.taskmap)
            if fkine is not None:
                if self.fkine is None:
                    self.fkine = fkine
                elif fkine is not self.fkine:
                    raise NotImplementedError("all FK task maps of one RmpCore must share one UrdfForwardKinematic")
            goal_slot = self.goal_leaves.index(name) if name in self.goal_leaves else -1
            if goal_slot >= 0 and space not in (_native.SPACE_FRAME_POSITION, _native.SPACE_FRAME_EULER):
                raise NotImplementedError("per-environment goals are implemented for frame position / orientation leaves")
            self.entries.append([leaf, space, frame, goal_slot, dist])
        if self.fkine is not None and self.fkine.n_joints != n:
            raise ValueError(f"q has {n} entries but the kinematics was built for {self.fkine.n_joints} joints")
        for e in self.entries:
            e[2] = self.fkine.frame_index(e[2]) if e[2] is not None else -1
        # leaves fed with explicit pair rows [K,8], in tree order
        self.pair_taskmaps = [e[4] for e in self.entries
                              if e[1] in (_native.SPACE_FRAME_DISTANCE_PAIRS, _native.SPACE_FRAME_POINTS)]
        self.pair_sources = [self._pair_source(e) for e in self.entries
                             if e[1] in (_native.SPACE_FRAME_DISTANCE_PAIRS, _native.SPACE_FRAME_POINTS)]
        self.uses_spheres = any(e[1] == _native.SPACE_FRAME_DISTANCE_SPHERES for e in self.entries)
        self._own_robot = None
        if self.fkine is not None:
            robot = self.fkine._handle
        else:                                                     # pure configuration-space tree
            robot = ctypes.c_void_p()
            T = np.eye(4, dtype=np.float32).reshape(1, 16)
            axis, jtype = np.zeros((1, 3), np.float32), np.zeros(1, np.int8)
            parent, qidx = np.full(1, -1, np.int32), np.full(1, -1, np.int32)     # alive until the call returns
            _native.check(_native.lib().rmp2_robot_create(
                T.ctypes.data, axis.ctypes.data, jtype.ctypes.data, parent.ctypes.data, qidx.ctypes.data, 1, n,
                ctypes.byref(robot)))
            self._own_robot = robot
        self.descs = self._make_descs()
        arr = (_native.LeafDesc * max(1, len(self.descs)))(*self.descs)
        self.handle = ctypes.c_void_p()
        _native.check(_native.lib().rmp2_tree_create(robot, arr, len(self.descs), ctypes.byref(self.handle)))

    @staticmethod
    def _pair_source(entry):
        """-> callable returning the leaf's current pair rows [K,8] (layout: include/rmp2_b200.h)."""
        leaf, space, _, _, taskmap = entry
        if space == _native.SPACE_FRAME_DISTANCE_PAIRS:
            return taskmap.current_pairs

        def rows():
            rel, data = taskmap.current_points(), leaf.current_data()
            if rel.shape[0] != data.shape[0]:
                raise ValueError(f"{leaf.name}: relative_pos has {rel.shape[0]} rows, d/vec have {data.shape[0]}")
            return torch.cat([rel, data, torch.zeros(rel.shape[0], 1)], dim=1)
        return rows

    def _make_desc(self, i):
        leaf, space, frame, goal_slot, _ = self.entries[i]
        return leaf.leaf_desc(self.n if space == _native.SPACE_CONFIG else 3, space, frame, goal_slot)

    def _make_descs(self):
        # (version read BEFORE the parameters: a concurrent assignment is then picked up by the next refresh)
        self._versions = [e[0].__dict__.get("_version", 0) for e in self.entries]
        self._vecs = {}                  # vector parameters as last seen by refresh (leaf index -> list of floats)
        return [self._make_desc(i) for i in range(len(self.entries))]

    def signature(self):
        return tuple((type(e[0]).__name__, e[1], e[2], e[3]) for e in self.entries)

    def refresh(self):
        """Push parameters the caller changed since the last step (e.g. ``target_rmp.goal = ...``,
        reference: experiments/franka_panda/06_cluttered_environment.py:142).  A leaf is looked at again when one of
        its attributes was assigned since (``_version``) or when it holds a vector parameter that is not fed per
        environment (goal / q0 / limits: arrays can be changed in place)."""
        for i, e in enumerate(self.entries):
            leaf = e[0]
            version = leaf.__dict__.get("_version", 0)
            if version == self._versions[i]:
                if e[3] >= 0 or not leaf._has_vector_parameters():
                    continue
                vec = leaf._vec(self.n if e[1] == _native.SPACE_CONFIG else 3)
                if vec == self._vecs.get(i):
                    continue
                self._vecs[i] = vec
            else:
                self._vecs.pop(i, None)
            d = self._make_desc(i)
            self._versions[i] = version
            if bytes(d) != bytes(self.descs[i]):
                _native.check(_native.lib().rmp2_tree_update_leaf(self.handle, i, d))
                self.descs[i] = d

    def __del__(self):
        try:
            _native.lib().rmp2_tree_destroy(self.handle)
            if self._own_robot is not None:
                _native.lib().rmp2_robot_destroy(self._own_robot)
        except Exception:
            pass

    # ---------------------------------------------------------------------------------------------
    def _io(self, B, q, qd, qdd, goals, spheres, pairs, pair_counts):
        io = _native.StepIO()
        io.B = B
        io.q, io.qd, io.qdd = q.data_ptr(), qd.data_ptr(), qdd.data_ptr()
        if goals is not None:
            io.goals, io.n_goal_slots = goals.data_ptr(), goals.shape[1]
        if spheres is not None and spheres.shape[1] > 0:
            io.spheres, io.n_spheres = spheres.data_ptr(), spheres.shape[1]
        io.n_pair_sets = len(pair_counts)
        for i, k in enumerate(pair_counts):
            io.pair_counts[i] = k
        if pairs is not None and pairs.shape[1] > 0:
            io.pairs = pairs.data_ptr()
        return io

    def _check(self, B, q, qd, qdd, goals, spheres, pairs, cuda):
        for name, t, shape in (("q", q, (B, self.n)), ("qd", qd, (B, self.n)), ("qdd", qdd, (B, self.n))):
            if tuple(t.shape) != shape or t.dtype != torch.float32 or not t.is_contiguous() or t.is_cuda != cuda:
                raise ValueError(f"{name} must be a contiguous float32 {'CUDA' if cuda else 'host'} tensor of shape {shape}")
        if self.goal_leaves:
            if goals is None or goals.dim() != 3 or goals.shape[0] != B or goals.shape[1] < len(self.goal_leaves) or goals.shape[2] != 3:
                raise ValueError(f"goals must be [B, {len(self.goal_leaves)}, 3]")
        for name, t, last in (("goals", goals, 3), ("spheres", spheres, 4), ("pairs", pairs, _native.PAIR_FLOATS)):
            if t is not None and (t.dim() != 3 or t.shape[0] != B or t.shape[2] != last or t.dtype != torch.float32
                                  or not t.is_contiguous() or t.is_cuda != cuda):
                raise ValueError(f"{name} must be a contiguous float32 {'CUDA' if cuda else 'host'} tensor [B, K, {last}]")

    def step(self, q, qd, qdd, goals=None, spheres=None, pairs=None, pair_counts=()):
        """Device tensors in, device tensor out; asynchronous on the current stream."""
        B = q.shape[0]
        self._check(B, q, qd, qdd, goals, spheres, pairs, cuda=True)
        if B == 0:
            return qdd
        io = self._io(B, q, qd, qdd, goals, spheres, pairs, pair_counts)
        _native.check(_native.lib().rmp2_step(self.handle, ctypes.byref(io), current_stream_ptr(q.device)))
        return qdd

    def step_host(self, q, qd, qdd, goals=None, spheres=None, pairs=None, pair_counts=()):
        """Host tensors in (pinned for full copy/compute overlap), host tensor out; returns when
        ``qdd`` is complete.  The library stages chunks through device memory it owns."""
        require_cuda()
        B = q.shape[0]
        self._check(B, q, qd, qdd, goals, spheres, pairs, cuda=False)
        io = self._io(B, q, qd, qdd, goals, spheres, pairs, pair_counts)
        _native.check(_native.lib().rmp2_step_host(self.handle, ctypes.byref(io)))
        return qdd

    def rollout(self, q, qd, qdd, dt, n_steps, control_every, goals=None, spheres=None):
        """In-place closed-loop rollout on the device (explicit Euler), see ``rmp2_rollout``."""
        B = q.shape[0]
        self._check(B, q, qd, qdd, goals, spheres, None, cuda=True)
        io = self._io(B, q, qd, qdd, goals, spheres, None, ())
        _native.check(_native.lib().rmp2_rollout(self.handle, ctypes.byref(io), q.data_ptr(), qd.data_ptr(),
                                                 float(dt), int(n_steps), int(control_every),
                                                 current_stream_ptr(q.device)))
        return q, qd, qdd

    KERNELS = ("frames", "spheres", "step", "resolve", "resolve_fallback")

    def kernel_info(self, n_spheres=64):
        """registers / shared memory / resident blocks per SM of the kernels one step can launch."""
        out = {}
        for which, name in enumerate(("frames", "spheres", "step_fused", "step", "resolve", "resolve_fallback")):
            if name in ("frames", "spheres") and (not self.uses_spheres or self.obstacle_slots()[1] == 0):
                continue                     # no pair loop runs (no obstacle leaves, or every one of them is inert)
            regs, smem, bps, block = (ctypes.c_int32() for _ in range(4))
            _native.check(_native.lib().rmp2_tree_kernel_info(self.handle, which, n_spheres, ctypes.byref(regs),
                                                              ctypes.byref(smem), ctypes.byref(bps), ctypes.byref(block)))
            out[name] = dict(registers=regs.value, smem_bytes=smem.value, blocks_per_sm=bps.value,
                             block_threads=block.value, warps_per_sm=bps.value * block.value // 32)
        return out

    def specialize(self, compile_only=False):
        """Rebuild the frames / step kernels for THIS tree with NVRTC (tables as compile-time constants, loops
        unrolled; a few seconds, once) and use them from now on -- worthwhile for large batches.  Changing a
        leaf parameter afterwards (``leaf.goal = ...``) drops the specialisation; call again to rebuild.
        ``compile_only`` runs the compiler without loading the result (needs no GPU).  Returns NVRTC seconds."""
        _native.check(_native.lib().rmp2_tree_specialize(self.handle, _native.SPECIALIZE_COMPILE_ONLY if compile_only else 0))
        return self.specialized_seconds()

    def specialized_seconds(self):
        """NVRTC compile time of the loaded specialisation, or None when the generic kernels are in use."""
        sec = ctypes.c_double()
        on = _native.lib().rmp2_tree_is_specialized(self.handle, ctypes.byref(sec))
        return sec.value if on else None

    def reserve(self, B, n_spheres=0):
        """Size the scratch for steps of up to B environments so that no later step allocates (needed before
        capturing steps in a CUDA graph)."""
        dev = require_cuda()
        _native.check(_native.lib().rmp2_tree_reserve(self.handle, int(B), int(n_spheres), current_stream_ptr(dev)))

    def set_option(self, option, value):
        _native.check(_native.lib().rmp2_tree_set_option(self.handle, int(option), int(value)))

    def set_early_out(self, enable=True):
        """Skip (frame, sphere) pairs beyond the metric radius in the obstacle kernel (exact; default on)."""
        _native.check(_native.lib().rmp2_tree_set_option(self.handle, _native.OPT_EARLY_OUT, 1 if enable else 0))

    def set_merge_coincident(self, enable=True):
        """Run one pair loop for obstacle leaves that share their control point (equal parameters, frame origins that
        coincide for every q; exact up to rounding; default on).  Off: every leaf on its own."""
        _native.check(_native.lib().rmp2_tree_set_option(self.handle, _native.OPT_MERGE_COINCIDENT, 1 if enable else 0))

    def obstacle_slots(self):
        """-> (sphere-path obstacle leaves of the tree, pair loops that run for them per environment)."""
        leaves, slots = ctypes.c_int32(), ctypes.c_int32()
        _native.check(_native.lib().rmp2_tree_obstacle_slots(self.handle, ctypes.byref(leaves), ctypes.byref(slots)))
        return leaves.value, slots.value

    def profile(self, enable=True):
        """Bracket every kernel launch of this tree with CUDA events (see ``profile_read``)."""
        _native.check(_native.lib().rmp2_tree_profile(self.handle, 1 if enable else 0))

    def profile_read(self):
        """-> {kernel: (milliseconds, launches)} accumulated since the last read."""
        ms = (ctypes.c_double * _native.PROFILE_KERNELS)()
        launches = (ctypes.c_int64 * _native.PROFILE_KERNELS)()
        _native.check(_native.lib().rmp2_tree_profile_read(self.handle, ms, launches))
        return {name: (ms[i], launches[i]) for i, name in enumerate(self.KERNELS)}


# =================================================================================================
# RmpCore
# =================================================================================================
class RmpCore:
    """Manages multiple RMPs and combines their commands into one (reference: rmp.py:111-155)."""

    def __init__(self, rmps=None):
        # the reference's default argument is one shared dict (rmp.py:114); a fresh dict per core is
        # what every experiment relies on in practice
        self.rmps = {} if rmps is None else rmps
        self._compiled = None
        self._compiled_key = None

    def __str__(self):
        if not self.rmps:
            return 'no RMPs in use.\n'
        out = '\nused RMPs:\n'
        for i, rmp in enumerate(self.rmps.values()):
            out += '\t'.join([str(i), rmp.name, str(type(rmp))]) + '\n'
        return out

    def add_rmp(self, rmp):
        self.rmps[rmp.name] = rmp

    def remove_rmp_by_name(self, name):
        self.rmps.pop(name)            # KeyError for unknown names, like the reference (rmp.py:131)

    # ---------------------------------------------------------------------------------------------
    def compile(self, n, goal_leaves=()):
        """Compile (or reuse) the kernel tables for the current leaf list."""
        key = (n, tuple(goal_leaves), tuple((name, id(leaf), id(leaf.taskmap)) for name, leaf in self.rmps.items()))
        if self._compiled is None or self._compiled_key != key:
            self._compiled = CompiledTree(self.rmps, n, goal_leaves)
            self._compiled_key = key
        else:
            self._compiled.refresh()
        return self._compiled

    def evaluate(self, q, qd, goals=None, spheres=None):
        """qdd = pinv(sum J^T M J) sum J^T M (xdd - Jdot qd)   (reference: rmp.py:133-155).

        q, qd: [n] as in the reference (returns [n]) or [B, n] (returns [B, n]).
        goals: optional per-environment goals, ``{leaf_name: [B,3]}`` or a single [B,3] array for the
               first TargetPolicy/TargetAttractor leaf.  spheres: optional [B, O, 4] = (centre, radius)
               for leaves on ``TaskmapJointFrame4x4ToSphereDistance``.
        NumPy / CPU inputs give a CPU tensor (``.numpy()`` works); CUDA inputs stay on the GPU.
        """
        dev = require_cuda()
        q_in = unwrap(q)
        single = (np.ndim(q_in) == 1) if not isinstance(q_in, torch.Tensor) else (q_in.dim() == 1)
        if not isinstance(goals, dict) and not any(is_device_tensor(unwrap(a)) for a in (q, qd, goals, spheres)):
            # all inputs live on the host (the reference's call): one staging copy for all of them
            q, qd, goals, spheres = stage_host_inputs((q, qd, goals, spheres), dev)
        qt = to_device(q, dev)
        qdt = to_device(qd, dev)
        qt = qt.reshape(1, -1) if single else qt
        qdt = qdt.reshape(1, -1) if single else qdt
        if qt.dim() != 2 or qt.shape != qdt.shape:
            raise ValueError("q and qd must both be [n] or [B, n]")
        B, n = qt.shape

        goal_names, goal_tensor = [], None
        if goals is not None:
            if not isinstance(goals, dict):
                first = next((name for name, leaf in self.rmps.items()
                              if isinstance(leaf, TargetPolicy) or type(leaf).__name__ == 'TargetAttractor'), None)
                if first is None:
                    raise ValueError("goals given but the tree has no target leaf")
                goals = {first: goals}
            goal_names = [name for name in self.rmps if name in goals]
            if len(goal_names) != len(goals):
                raise KeyError(f"goals for unknown leaves: {set(goals) - set(goal_names)}")
            goal_tensor = torch.stack([to_device(goals[name], dev).reshape(B, 3) for name in goal_names], dim=1).contiguous()
        tree = self.compile(n, goal_names)

        sph = None
        if spheres is not None:
            sph = to_device(spheres, dev)
            if sph.dim() == 2:
                sph = sph[None]
        pairs, counts = None, []
        if tree.pair_taskmaps:
            per_leaf = [rows() for rows in tree.pair_sources]
            counts = [int(p.shape[0]) for p in per_leaf]
            if B != 1:
                raise NotImplementedError("explicit closest-point pairs (Datamanager feed) describe one environment; "
                                          "use spheres=[B,O,4] for batched obstacles")
            if sum(counts) > 0:
                pairs = torch.cat([p.to(dev) for p in per_leaf], dim=0)[None].contiguous()
        qdd = torch.empty(B, n, device=dev, dtype=torch.float32)
        tree.step(qt, qdt, qdd, goals=goal_tensor, spheres=sph, pairs=pairs, pair_counts=counts)
        out = qdd[0] if single else qdd
        return out if is_device_tensor(q_in) else out.cpu()
