"""The C-ABI shared library: builds, loads and exports every symbol include/rmp2_b200.h declares;
host-only entry points (handle creation, tree compilation, argument checking) work without a GPU.
No compute call is made here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from riemannian_motion_policies_b200 import _native, scenarios as S


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rmp2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rmp2_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol(native_lib):
    for sym in _declared_symbols():
        assert hasattr(native_lib, sym), sym
    assert b"sm_100a" in native_lib.rmp2_version()


def test_struct_layouts_match_header():
    # rmp2_leaf_desc: 4 int32 + 16 float + 24 float ; rmp2_step_io as declared
    assert ctypes.sizeof(_native.LeafDesc) == 4 * 4 + 4 * _native.RMP2_LEAF_PARAMS + 4 * 2 * _native.RMP2_MAX_JOINTS
    io = _native.StepIO
    assert io.B.offset == 0 and io.q.offset == 8 and io.qdd.offset == 24 and io.goals.offset == 32
    assert io.n_goal_slots.offset == 40 and io.n_spheres.offset == 44 and io.spheres.offset == 48
    assert io.pairs.offset == 56 and io.n_pair_sets.offset == 64 and io.pair_counts.offset == 68


def _robot(native_lib, F=3, n=2, parent=(-1, 0, 1), jtype=(1, 1, 0), axis=None):
    T = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (F, 1))
    axis = np.tile(np.array([[0, 0, 1]], np.float32), (F, 1)) if axis is None else np.asarray(axis, np.float32)
    jt = np.asarray(jtype, np.int8)
    par = np.asarray(parent, np.int32)
    qidx = np.array([0, 1, -1][:F], np.int32)
    h = ctypes.c_void_p()
    rc = native_lib.rmp2_robot_create(T.ctypes.data, axis.ctypes.data, jt.ctypes.data, par.ctypes.data,
                                      qidx.ctypes.data, F, n, ctypes.byref(h))
    return rc, h


def test_robot_create_checks_arguments(native_lib):
    rc, h = _robot(native_lib)
    assert rc == 0 and h.value
    native_lib.rmp2_robot_destroy(h)
    rc, _ = _robot(native_lib, parent=(-1, 2, 1))
    assert rc == 1 and b"parent" in native_lib.rmp2_last_error()
    rc, _ = _robot(native_lib, axis=[[0, 0, 2], [0, 0, 1], [0, 0, 0]])
    assert rc == 3 and b"unit length" in native_lib.rmp2_last_error()


def test_tree_create_rejects_unsupported_combinations(native_lib):
    rc, h = _robot(native_lib)
    d = _native.LeafDesc()
    d.type, d.space, d.frame, d.goal_slot = _native.LEAF_OBSTACLE_AVOIDANCE, _native.SPACE_CONFIG, -1, -1
    t = ctypes.c_void_p()
    rc = native_lib.rmp2_tree_create(h, ctypes.byref(d), 1, ctypes.byref(t))
    assert rc == 3
    d.type, d.space, d.frame = _native.LEAF_TARGET_ATTRACTOR, _native.SPACE_FRAME_POSITION, 7
    rc = native_lib.rmp2_tree_create(h, ctypes.byref(d), 1, ctypes.byref(t))
    assert rc == 1 and b"frame index" in native_lib.rmp2_last_error()
    native_lib.rmp2_robot_destroy(h)


def test_python_tree_compiler_on_cpu(native_lib):
    """Tree compilation is host-only: every workload of BASELINE.json compiles without a GPU, the
    unsupported chains raise NotImplementedError (no fallback)."""
    ns = S.product_namespace()
    fk = ns.UrdfForwardKinematic(S.PANDA_URDF, S.PANDA_ORDER_9)
    assert fk.frame_names[-1] == "panda_grasptarget_hand"
    sphere_tm = lambda frame: ns.TaskmapJointFrame4x4ToSphereDistance()
    for cfg in (3, 4, 5):
        core = S.BUILDERS[cfg](ns, fk, [0.5, 0.0, 0.5], 9, sphere_tm)
        tree = core.compile(9, goal_leaves=["attractor"])
        assert tree.uses_spheres and tree.handle.value
    core = S.build_config2(ns, fk, [0.5, 0.0, 0.5], 9)
    assert not core.compile(9).uses_spheres
    # explicit-pair leaves (Datamanager feed)
    dm = ns.Datamanager(fk)
    core = S.build_config3(ns, fk, [0.5, 0, 0.5], 9, lambda fr: ns.TaskmapJointFrame4x4ToDistance(
        dm[fr]['pos_on_link_in_base_frame'], dm[fr]['pos_on_obstacle_in_base_frame']))
    assert len(core.compile(9).pair_taskmaps) == 10
    # a user-defined task map cannot be compiled
    core = ns.RmpCore()
    core.add_rmp(ns.TargetPolicy(0.1, 1, 0.1, [0, 0, 0], ns.TaskmapByFunction(lambda q: q, lambda q, qd: None)))
    with pytest.raises(NotImplementedError):
        core.compile(9)
    core = ns.RmpCore()
    core.add_rmp(ns.CollisionAvoidance(None, None, 1, 1, 1, 1, 1, 1, ns.IdentityTaskmap()))
    with pytest.raises(NotImplementedError):
        core.compile(9)
    with pytest.raises(KeyError):
        ns.RmpCore().remove_rmp_by_name("missing")


def test_compute_fails_loudly_without_cuda(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    ns = S.product_namespace()
    fk = ns.UrdfForwardKinematic(S.TWO_JOINT_URDF, S.TWO_JOINT_ORDER)
    core = S.build_config1(ns, fk, [1.0, 0.5, 0.1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        core.evaluate(np.zeros(2, np.float32), np.zeros(2, np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fk.forward(np.zeros((1, 2), np.float32), "link_23")


def test_specialization_compiles_without_a_gpu():
    """rmp2_tree_specialize(COMPILE_ONLY): NVRTC builds the tree-specialised frames / step kernels from the
    sources embedded in the library (no GPU needed; loading and launching them is covered by the gpu tests)."""
    from riemannian_motion_policies_b200 import scenarios as S
    ns = S.product_namespace()
    for urdf, order, n, build in ((S.PANDA_WO_TOOL_URDF, S.PANDA_ORDER_7, 7, S.build_config5),
                                  (S.PANDA_URDF, S.PANDA_ORDER_9, 9, S.build_config3)):
        fk = ns.UrdfForwardKinematic(urdf, order)
        core = build(ns, fk, [0.5, 0.0, 0.5], n, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
        tree = core.compile(n, goal_leaves=["attractor"])
        assert tree.specialize(compile_only=True) is None        # compiled, nothing loaded


def test_gantry_tree_with_orientation_leaf_compiles_and_specializes(native_lib):
    """The synthetic gantry arm (general axes, three branchings, nine scrambled joints) with an orientation leaf on
    the Euler task map (RMP2_SPACE_FRAME_EULER): host-side compilation and the NVRTC build need no GPU."""
    ns = S.product_namespace()
    fk = ns.UrdfForwardKinematic(S.GANTRY_URDF, S.GANTRY_ORDER)
    core = S.build_config6(ns, fk, [0.2, 0.1, 0.5], 9, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
    tree = core.compile(9, goal_leaves=["target"])
    spaces = [d.space for d in tree.descs]
    assert _native.SPACE_FRAME_EULER in spaces and _native.SPACE_FRAME_POSITION in spaces
    assert tree.specialize(compile_only=True) is None
    # an orientation leaf accepts only the two target policies
    bad = ns.RmpCore()
    bad.add_rmp(ns.JointDamping(accel_d_gain=1, metric_scalar=0.005, inertia=0.3))
    bad.rmps["joint_damping"].taskmap = S.euler_taskmap(ns, fk, "tool")
    with pytest.raises(NotImplementedError):
        bad.compile(9)


def test_step_argument_checks_need_no_gpu(native_lib):
    """Everything rmp2_step refuses is refused before any CUDA call: misaligned sphere rows, missing goals, pair-set
    mismatches; likewise the options, rmp2_tree_reserve and rmp2_pinv_solve argument checks."""
    ns = S.product_namespace()
    fk = ns.UrdfForwardKinematic(S.PANDA_WO_TOOL_URDF, S.PANDA_ORDER_7)
    core = S.build_config4(ns, fk, [0.5, 0.0, 0.5], 7, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
    tree = core.compile(7, goal_leaves=["attractor"])
    io = _native.StepIO()
    io.B, io.q, io.qd, io.qdd = 4, 0x1000, 0x2000, 0x3000          # never dereferenced: the call fails first
    io.goals, io.n_goal_slots = 0x4000, 1
    io.spheres, io.n_spheres = 0x5004, 8                           # 4-byte aligned only
    assert native_lib.rmp2_step(tree.handle, ctypes.byref(io), None) == 1
    assert b"16-byte aligned" in native_lib.rmp2_last_error()
    io.spheres, io.goals = 0x5000, None
    assert native_lib.rmp2_step(tree.handle, ctypes.byref(io), None) == 1
    assert b"per-environment goals" in native_lib.rmp2_last_error()
    io.goals, io.n_pair_sets = 0x4000, 2
    assert native_lib.rmp2_step(tree.handle, ctypes.byref(io), None) == 1
    assert b"n_pair_sets" in native_lib.rmp2_last_error()
    assert native_lib.rmp2_rollout(tree.handle, None, 0x1000, 0x2000, 0.01, 10, 10, None) == 1   # NULL io: no crash
    for option, value in ((_native.OPT_SPLIT_RESOLVE, 2), (_native.OPT_BLOCK_THREADS, 48), (_native.OPT_CHUNK_ENVS, -1), (99, 0)):
        assert native_lib.rmp2_tree_set_option(tree.handle, option, value) == 1
    for option, value in ((_native.OPT_SPLIT_RESOLVE, 0), (_native.OPT_BLOCK_THREADS, 64), (_native.OPT_TMA, 0),
                          (_native.OPT_EARLY_OUT, 0), (_native.OPT_CHUNK_ENVS, 1 << 16)):
        assert native_lib.rmp2_tree_set_option(tree.handle, option, value) == 0
    assert native_lib.rmp2_tree_reserve(None, 16, 0, None) == 1
    assert native_lib.rmp2_tree_reserve(tree.handle, -1, 0, None) == 1
    assert native_lib.rmp2_pinv_solve(7, 4, None, 0x1000, 0x2000, 1, 0, None) == 1
    assert native_lib.rmp2_pinv_solve(13, 4, 0x1000, 0x2000, 0x3000, 1, 0, None) == 1
    assert native_lib.rmp2_pinv_solve(7, 4, 0x1000, 0x2000, 0x3000, 1, 2, None) == 1
    assert native_lib.rmp2_obstacle_feed(fk._handle, np.zeros(1, np.int32).ctypes.data, None, 1, 4, 0x1000, 0x5004, 2, None, 0,
                                         0x6000, None, None) == 1
    assert b"16-byte aligned" in native_lib.rmp2_last_error()


def test_coincident_obstacle_leaves_are_merged(native_lib):
    """RMP2_OPT_MERGE_COINCIDENT: obstacle leaves with equal parameters on frames whose origins coincide for every q run
    one pair loop (Panda: joint6 sits on joint5; the finger joints are prismatic and stay), and leaves whose control
    point cannot move at all (joint1 on the base axis, joint2 on top of it: J = 0, their pulled-back (M, f) is exactly
    zero in the reference too) run none.  A leaf whose parameters differ leaves its group; switching the option off
    gives every leaf its own loop again."""
    ns = S.product_namespace()
    dist = lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance()
    fk7 = ns.UrdfForwardKinematic(S.PANDA_WO_TOOL_URDF, S.PANDA_ORDER_7)
    core = S.build_config4(ns, fk7, [0.5, 0.0, 0.5], 7, dist)
    tree = core.compile(7)
    assert tree.obstacle_slots() == (8, 5)                      # {joint1, joint2} inert, {joint5, joint6} merged
    assert tree.specialize(compile_only=True) is None           # the merged tables still compile under NVRTC
    tree.set_merge_coincident(False)
    assert tree.obstacle_slots() == (8, 8)
    tree.set_merge_coincident(True)
    assert tree.obstacle_slots() == (8, 5)
    # the reference idiom `leaf.attr = value`, picked up at the next step: joint6's leaf leaves joint5's group
    core.rmps["collision_avoidance_for_panda_joint6"].repulsion_gain = 700.
    tree = core.compile(7)
    assert tree.obstacle_slots() == (8, 6)
    core.rmps["collision_avoidance_for_panda_joint6"].repulsion_gain = 800.
    assert core.compile(7).obstacle_slots() == (8, 5)
    fk9 = ns.UrdfForwardKinematic(S.PANDA_URDF, S.PANDA_ORDER_9)
    assert S.build_config4(ns, fk9, [0.5, 0.0, 0.5], 9, dist).compile(9).obstacle_slots() == (10, 7)
    assert native_lib.rmp2_tree_obstacle_slots(None, None, None) == 1
    # a tree whose only obstacle leaves are inert still compiles (no pair loop at all)
    inert = ns.RmpCore()
    inert.add_rmp(S.target_attractor(ns, fk7, [0.5, 0.0, 0.5]))
    for fr in ("panda_joint1", "panda_joint2"):
        inert.add_rmp(S.obstacle_leaf(ns, ns.chain_taskmaps([ns.TaskmapByForwardKinematic(fk7, fr), dist(fr)]), fr))
    tree = inert.compile(7)
    assert tree.obstacle_slots() == (2, 0) and tree.specialize(compile_only=True) is None


def test_refresh_picks_up_leaf_changes(native_lib):
    """The reference's callers change leaf parameters by assignment (``target_rmp.goal = ...``,
    06_cluttered_environment.py:142) or mutate a goal array in place; a compiled tree re-derives exactly the leaves that
    moved (attribute version counter; vector parameters compared by value) and pushes them with rmp2_tree_update_leaf."""
    ns = S.product_namespace()
    fk = ns.UrdfForwardKinematic(S.PANDA_WO_TOOL_URDF, S.PANDA_ORDER_7)
    core = S.build_config4(ns, fk, [0.5, 0.0, 0.5], 7, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
    tree = core.compile(7)
    names = list(core.rmps)
    before = [bytes(d) for d in tree.descs]
    assert core.compile(7) is tree and [bytes(d) for d in tree.descs] == before        # nothing changed
    goal = np.array([0.5, 0.0, 0.5])
    core.rmps["attractor"].goal = goal                                                 # same values: no update
    core.compile(7)
    assert [bytes(d) for d in tree.descs] == before
    goal[1] = 0.25                                                                     # in place
    core.compile(7)
    i = names.index("attractor")
    assert list(tree.descs[i].vec[:3]) == [0.5, 0.25, 0.5]
    core.rmps["joint_limit_avoidance"].gamma_p = 0.4                                   # scalar by assignment
    core.compile(7)
    j = names.index("joint_limit_avoidance")
    assert abs(tree.descs[j].params[0] - 0.4) < 1e-7
    changed = [k for k, d in enumerate(tree.descs) if bytes(d) != before[k]]
    assert sorted(changed) == sorted([i, j])
    core.rmps["attractor"].goal = [0.5, 0.0]                                           # wrong length is refused
    with pytest.raises(ValueError):
        core.compile(7)


def test_host_inputs_travel_in_one_staging_buffer():
    """_tensor.stage_host_inputs: every host argument of a call becomes a float32 view of ONE buffer (one host-to-device
    copy), sections on 16-byte boundaries (sphere rows are read as float4), values rounded exactly like to_device."""
    import torch
    from riemannian_motion_policies_b200._tensor import stage_host_inputs
    q = np.linspace(-1, 1, 7)                                   # float64 ndarray
    qd = [0.1 * i for i in range(7)]                            # list
    sph = torch.arange(24, dtype=torch.float64).reshape(2, 3, 4) / 7
    out = stage_host_inputs((q, qd, None, sph), torch.device("cpu"))
    assert out[2] is None and [tuple(o.shape) for o in out if o is not None] == [(7,), (7,), (2, 3, 4)]
    assert all(o.dtype == torch.float32 and o.is_contiguous() for o in out if o is not None)
    base = out[0].data_ptr()
    assert [(o.data_ptr() - base) % 16 for o in out if o is not None] == [0, 0, 0]
    assert (out[1].data_ptr() - base, out[3].data_ptr() - base) == (32, 64)      # 7 floats padded to 8
    np.testing.assert_array_equal(out[0].numpy(), q.astype(np.float32))
    np.testing.assert_array_equal(out[3].numpy(), sph.to(torch.float32).numpy())


def test_oracle_sensitivity_yardstick():
    """oracle/harness.config_sensitivity: deterministic per environment (independent of the batch around it), at the
    eps32 scale for a well-conditioned tree, and kappa * eps32 for a nearly singular metric."""
    import torch
    from oracle import harness as H
    q, qd, goal = S.sample_panda_state(6, 7, seed=3)
    full = H.config_sensitivity(2, 7, q, qd, goal)
    part = H.config_sensitivity(2, 7, q[2:4], qd[2:4], goal[2:4])
    np.testing.assert_array_equal(full[2:4], part)
    assert (full > 1e-8).all() and (full < 1e-4).all()
    M = np.diag([1.0, 1e-3, 1e-9, 0, 0, 0, 0])[None]
    np.testing.assert_allclose(H.metric_conditioning(M), [1e3 * H.EPS32])       # 1e-9 is below the pinv cutoff
