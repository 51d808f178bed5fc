"""The trees of the reference's experiment scripts rebuilt from `scenarios` builders -- the form that travels to the
GPU box (the scripts themselves do not).  tests/test_dropin_experiments.py proves, in the build container, that
exec'ing the scripts' own lines on top of compat/ yields exactly these leaf descriptors."""
import importlib
import os
import sys
import types

import numpy as np

from riemannian_motion_policies_b200 import scenarios as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "riemannian_motion_policies_b200", "compat")
GOALS = {"exp06": [0.2, -0.2, 0.5], "exp04": [0.6, 0, 0.5], "two01": [1.4, -1.4, 0.1], "two05": [1.4, -1.4, 0.1]}
ROBOT = {"exp06": "panda", "exp04": "panda", "two01": "two_joint", "two03": "two_joint", "two05": "two_joint"}


def compat_namespace():
    """The product's classes imported the way a reference script imports them: `from rmp import ...` with
    riemannian_motion_policies_b200/compat first on sys.path."""
    names = ("kinematics", "taskmap", "rmp", "rmp2", "data_management")
    saved_path, saved = list(sys.path), {k: sys.modules.pop(k, None) for k in names}
    sys.path.insert(0, COMPAT)
    try:
        ns = types.SimpleNamespace()
        for name in names:
            mod = importlib.import_module(name)
            assert os.path.dirname(os.path.abspath(mod.__file__)) == COMPAT, mod.__file__
            for k, v in vars(mod).items():
                if not k.startswith("_"):
                    setattr(ns, k, v)
        return ns
    finally:
        sys.path[:] = saved_path
        for k in names:
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]


def make_fkine(ns, key):
    if ROBOT[key] == "panda":
        return ns.UrdfForwardKinematic(S.PANDA_URDF, S.PANDA_ORDER_9)
    return ns.UrdfForwardKinematic(S.TWO_JOINT_URDF, S.TWO_JOINT_ORDER)


def build(ns, key, fk, dm=None):
    """-> RmpCore equal to what the script's block builds.  `dm`: a Datamanager-like mapping frame -> variables."""
    if key == "exp06":                                            # 06_cluttered_environment.py, block "# forward kinematic"
        return S.build_config3(ns, fk, GOALS[key], 9, lambda fr: ns.TaskmapJointFrame4x4ToDistance(
            pos_on_link_in_base_frame=dm[fr]['pos_on_link_in_base_frame'],
            pos_on_obstacle_in_base_frame=dm[fr]['pos_on_obstacle_in_base_frame']))
    if key == "exp04":                                            # 04_nullspace_control.py
        return S.build_config2(ns, fk, GOALS[key], 9)
    if key == "two01":                                            # two_joint_robot/01_target_rmp_only.py
        return S.build_config1(ns, fk, GOALS[key])
    if key == "two03":                                            # two_joint_robot/03_jointlimit_avoiding.py
        core = ns.RmpCore()
        core.add_rmp(ns.JointLimitAvoidance(np.array([-np.pi, -np.pi]), np.array([np.pi, np.pi]), gamma_p=0.3, gamma_d=1))
        return core
    if key == "two05":                                            # two_joint_robot/05_obstacle_avoidance.py
        core = ns.RmpCore()
        core.add_rmp(ns.TargetPolicy(alpha=0.1, beta=0.1, c=0.1, goal=GOALS[key], name='target',
                                     taskmap=S.ee_position_taskmap(ns, fk, 'link_23')))
        for frame in fk.frame_names:
            tm = ns.chain_taskmaps([ns.TaskmapByForwardKinematic(fk, frame),
                                    ns.TaskmapRelative4x4(relative_pos=dm[frame]['relative_position']),
                                    ns.TaskmapFrom4x4ToPosition()])
            core.add_rmp(ns.CollisionAvoidance(d=dm[frame]['distance'], vec=dm[frame]['normal_vec'],
                                               eta_rep=0.1 * np.e, nu_rep=0.3, eta_damp=1, nu_damp=0.3, r=1.1, c=1e5,
                                               taskmap=tm, name=f'collision_avoidance_for_{frame}'))
        return core
    raise KeyError(key)


def unpack_rows(rows, frames):
    """float32 rows [K,11] of ref_exp_*.npz -> the reference's distance_data tuples (simulation.py:462-484)."""
    return [(str(frames[int(r[0])]), r[1:4].copy(), r[4:7].copy(), r[7:10].copy(), np.float32(r[10]), "synthetic")
            for r in rows]


def oracle_evaluate(key, g, b, dtype):
    """The oracle on environment b of fixture g (same tuples; Datamanager restated as plain tensors)."""
    import torch
    from oracle import harness as H
    ons = H.namespace(dtype)
    fk = H.make_fkine(9 if ROBOT[key] == "panda" else 2, dtype)
    lo, hi = int(g["row_count"][:b].sum()), int(g["row_count"][:b + 1].sum())
    data = unpack_rows(g["rows"][lo:hi], list(g["frames"]))
    q, qd = torch.as_tensor(g["q"][b]), torch.as_tensor(g["qd"][b])
    dm = {}
    for frame in fk.frame_names:
        rows = [d for d in data if d[0] == frame]
        T = fk.forward(q[None], frame)[0]
        rel = [T[:3, :3].T @ (torch.as_tensor(d[1]).to(dtype) - T[:3, 3]) for d in rows]     # data_management.py:44-52
        as3 = lambda k: torch.tensor(np.array([d[k] for d in rows]).reshape(-1, 3), dtype=dtype)
        dm[frame] = {'pos_on_link_in_base_frame': as3(1), 'pos_on_obstacle_in_base_frame': as3(2), 'normal_vec': as3(3),
                     'distance': torch.tensor([float(d[4]) for d in rows], dtype=dtype),
                     'relative_position': torch.stack(rel) if rel else torch.zeros(0, 3, dtype=dtype)}
    core = build(ons, key, fk, dm)
    return core.evaluate(q, qd).numpy()
