"""Run the REFERENCE'S OWN source files on seeded inputs and save inputs + outputs as golden vectors.

    python tests/golden/run_reference_under_shim.py [out_dir]

Build container only (needs /root/reference).  TensorFlow and PyBullet are not installed, so the reference
modules (`kinematics.py`, `taskmap.py`, `rmp.py`, `rmp2.py`, `data_management.py`, `helper/rmp_helper.py`) are
imported UNCHANGED with `oracle/tf_shim` on sys.path: a minimal TensorFlow-API stand-in over torch plus empty
pybullet modules (see oracle/tf_shim/README.md).  The trees are built by the same scenario builders the tests
use, with the reference's classes as the namespace.  Output: tests/golden/ref_*.npz.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, REFERENCE)

import tensorflow as tf                                            # noqa: E402  (the shim)
import torch                                                       # noqa: E402
import data_management                                             # noqa: E402  (reference modules)
import kinematics                                                  # noqa: E402
import rmp                                                         # noqa: E402
import rmp2                                                        # noqa: E402
import taskmap                                                     # noqa: E402
from oracle import harness as H                                    # noqa: E402
from riemannian_motion_policies_b200 import scenarios as S         # noqa: E402

assert "tf_shim" in tf.__file__ and kinematics.__file__.startswith(REFERENCE)


def reference_namespace():
    ns = types.SimpleNamespace()
    for mod in (kinematics, taskmap, rmp, rmp2, data_management):
        for k, v in vars(mod).items():
            if not k.startswith("_"):
                setattr(ns, k, v)
    core_cls = ns.RmpCore
    ns.RmpCore = lambda: core_cls(rmps={})        # the reference's default dict is shared between instances
    return ns


URDFS = {2: ("urdf/TwoJointRobot_wo_fixedJoints.urdf", S.TWO_JOINT_ORDER),
         7: ("urdf/franka_panda/panda_wo_tool.urdf", S.PANDA_ORDER_7),
         9: ("urdf/franka_panda/panda.urdf", S.PANDA_ORDER_9)}


def run_config(ns, config, n, B):
    if config == 6:
        return run_gantry(ns, B)
    path, order = URDFS[n]
    fk = ns.UrdfForwardKinematic(os.path.join(REFERENCE, path), order)
    ofk = H.make_fkine(n, torch.float64)
    fk.has_collision = ofk.has_collision          # the experiments ask PyBullet; same information
    seed = S.SEEDS[config] + 50                   # distinct from the oracle-generated fixtures
    if config == 1:
        q, qd, goal = S.sample_two_joint(B, seed)
    else:
        q, qd, goal = S.sample_panda_state(B, n, seed)
    O_ = 0 if config == 1 else S.N_SPHERES[config]
    frames = S.collision_frames(ofk)
    spheres = np.zeros((B, max(O_, 1), 4), np.float32)
    out = []
    n_eval = min(B, LIMIT) if LIMIT else B          # --limit: the first environments of the SAME seeded input set
    for b in range(n_eval):
        if O_:
            origins = H.frame_origins(ofk, torch.as_tensor(q[b]).double(), frames).numpy()
            sph = S.sample_spheres(1, O_, seed + b, origins[None])[0]
            spheres[b] = sph
            on_link, on_obst = S.closest_points_on_spheres(origins, sph)
            idx = {fr: i for i, fr in enumerate(frames)}
            tm_for = lambda fr: ns.TaskmapJointFrame4x4ToDistance(tf.constant(on_link[idx[fr]]), tf.constant(on_obst[idx[fr]]))
            core = S.BUILDERS[config](ns, fk, goal[b], n, tm_for)
        elif config == 1:
            core = S.build_config1(ns, fk, goal[b])
        else:
            core = S.build_config2(ns, fk, goal[b], n)
        res = core.evaluate(q[b], qd[b])
        assert res.numpy().dtype == np.float32      # accumulators become float32 tensors
        out.append(res.numpy())
    d = dict(q=q[:n_eval], qd=qd[:n_eval], goal=goal[:n_eval], qdd_ref=np.stack(out))
    if O_:
        d["spheres"] = spheres[:n_eval]
    return d


def run_gantry(ns, B):
    """config 6: this repo's synthetic gantry arm (general axes, x/y/z prismatic joints, multi-axis rpy, three
    branchings) read by the reference's own URDF parser, with an ORIENTATION leaf on the reference's
    TaskmapFrom4x4ToEuler (taskmap.py:57-67) -- scenarios.build_config6 with the reference's classes."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gpu_common import make_inputs
    fk = ns.UrdfForwardKinematic(S.GANTRY_URDF, S.GANTRY_ORDER)
    ofk = H.make_fkine(9, torch.float64, robot="gantry")
    fk.has_collision = ofk.has_collision
    frames = S.collision_frames(ofk)
    q, qd, goal, spheres = make_inputs(6, 9, B, seed=S.SEEDS[6] + 50)
    out = []
    B = min(B, LIMIT) if LIMIT else B
    q, qd, goal, spheres = q[:B], qd[:B], goal[:B], spheres[:B]
    for b in range(B):
        origins = H.frame_origins(ofk, torch.as_tensor(q[b]).double(), frames).numpy()
        on_link, on_obst = S.closest_points_on_spheres(origins, spheres[b])
        idx = {fr: i for i, fr in enumerate(frames)}
        tm_for = lambda fr: ns.TaskmapJointFrame4x4ToDistance(tf.constant(on_link[idx[fr]]), tf.constant(on_obst[idx[fr]]))
        core = S.build_config6(ns, fk, goal[b], 9, tm_for)
        out.append(core.evaluate(q[b], qd[b]).numpy())
    return dict(q=q, qd=qd, goal=goal, spheres=spheres, qdd_ref=np.stack(out))


def run_v1_two_joint(ns, B):
    """experiments/two_joint_robot/05_obstacle_avoidance.py:44-61 with synthetic distance_data."""
    path, order = URDFS[2]
    fk = ns.UrdfForwardKinematic(os.path.join(REFERENCE, path), order)
    rng = np.random.RandomState(77)
    q_all, qd_all, goal_all = S.sample_two_joint(B, seed=78)
    rows, outs = [], []
    B = min(B, LIMIT) if LIMIT else B
    q_all, qd_all, goal_all = q_all[:B], qd_all[:B], goal_all[:B]
    for b in range(B):
        q, qd, goal = q_all[b], qd_all[b], goal_all[b]
        distance_data = []
        for frame in fk.frame_names:
            for _ in range(2):
                T = fk.forward(tf.constant([q], dtype=tf.float32), tf.constant(frame)).numpy()[0]
                on_link = T[:3, 3] + T[:3, :3] @ rng.uniform(-0.3, 0.3, size=3)
                direction = rng.normal(size=3)
                direction /= np.linalg.norm(direction)
                dist = rng.uniform(0.05, 1.3)
                distance_data.append((frame, on_link.astype(np.float32), (on_link - dist * direction).astype(np.float32),
                                      direction.astype(np.float32), np.float32(dist), "synthetic"))
        dm = ns.Datamanager(fk)
        core = ns.RmpCore()
        core.add_rmp(ns.TargetPolicy(alpha=0.1, beta=0.1, c=0.1, goal=goal, name="target",
                                     taskmap=S.ee_position_taskmap(ns, fk, "link_23")))
        for frame in fk.frame_names:
            tm = ns.chain_taskmaps([ns.TaskmapByForwardKinematic(fk, frame),
                                    ns.TaskmapRelative4x4(relative_pos=dm[frame]["relative_position"]),
                                    ns.TaskmapFrom4x4ToPosition()])
            core.add_rmp(ns.CollisionAvoidance(d=dm[frame]["distance"], vec=dm[frame]["normal_vec"],
                                               eta_rep=0.1 * np.e, nu_rep=0.3, eta_damp=1, nu_damp=0.3, r=1.1, c=1e5,
                                               taskmap=tm, name=f"collision_avoidance_for_{frame}"))
        dm.update(q, distance_data)
        outs.append(core.evaluate(q, qd).numpy())
        rows.append(np.array([np.concatenate([d[1], d[2], d[3], [d[4]]]) for d in distance_data], dtype=np.float32))
    return dict(q=q_all, qd=qd_all, goal=goal_all, qdd_ref=np.stack(outs), distance_rows=np.stack(rows),
                frames=np.array([d[0] for d in distance_data]))


def run_fk(ns, n, B):
    path, order = URDFS[n]
    fk = ns.UrdfForwardKinematic(os.path.join(REFERENCE, path), order)
    rng = np.random.RandomState(200 + n)
    lo, hi = (S.PANDA_Q_LOW[:n], S.PANDA_Q_HIGH[:n]) if n > 2 else (-np.pi * np.ones(2), np.pi * np.ones(2))
    q = rng.uniform(lo, hi, size=(B, n)).astype(np.float32)
    qd = rng.uniform(-1, 1, size=(B, n)).astype(np.float32)
    d = dict(q=q, qd=qd, frame_names=np.array(fk.frame_names))
    for fi, frame in enumerate(fk.frame_names):
        res = [fk.differentiate(tf.constant([q[b]], dtype=tf.float32), tf.constant([qd[b]], dtype=tf.float32),
                                tf.constant(frame)) for b in range(B)]
        for k, name in enumerate(("x", "xd", "J", "c")):
            d[f"{name}_{fi}"] = np.stack([r[k].numpy()[0] for r in res])
    return d


LIMIT = 0
CONFIG_CASES = ((1, 2, 16), (2, 7, 12), (2, 9, 6), (3, 7, 8), (3, 9, 4), (4, 7, 256), (5, 7, 256), (6, 9, 12))


def main():
    """usage: run_reference_under_shim.py [out_dir] [--limit K] [--only NAME ...]   (NAME e.g. ref_config4_n7; the two
    256-environment cases take 10-30 s per environment -- the shim's batch_jacobian loops over the 64 pairs of a
    frame -- so they can be run side by side; --limit K evaluates only the first K environments of each case)"""
    global LIMIT
    argv = sys.argv[1:]
    if "--limit" in argv:
        k = argv.index("--limit")
        LIMIT = int(argv[k + 1])
        del argv[k:k + 2]
    only = None
    if "--only" in argv:
        k = argv.index("--only")
        only, argv = set(argv[k + 1:]), argv[:k]
    out_dir = argv[0] if argv else HERE
    want = lambda name: only is None or name[:-4] in only
    ns = reference_namespace()
    import builtins
    real_print, builtins.print = builtins.print, lambda *a, **k: None      # the reference prints on every FK build
    try:
        results = {f"ref_config{c}_n{n}.npz": run_config(ns, c, n, B)
                   for c, n, B in CONFIG_CASES if want(f"ref_config{c}_n{n}.npz")}
        if want("ref_v1_two_joint.npz"):
            results["ref_v1_two_joint.npz"] = run_v1_two_joint(ns, 8)
        for n, B in ((2, 4), (9, 3)):
            if want(f"ref_fk_n{n}.npz"):
                results[f"ref_fk_n{n}.npz"] = run_fk(ns, n, B)
    finally:
        builtins.print = real_print
    for name, d in results.items():
        np.savez_compressed(os.path.join(out_dir, name), **d)
        print("wrote", name)


if __name__ == "__main__":
    main()
