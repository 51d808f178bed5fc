"""Workloads of BASELINE.json / SURVEY.md section 8d: tree builders and seeded synthetic inputs.

Every builder takes a *namespace* ``ns`` that provides the reference's class names
(``RmpCore``, ``TargetAttractor``, ``chain_taskmaps`` ...).  The product namespace is
``product_namespace()``; the tests pass the CPU oracle module instead, so both sides are built by
the very same lines, with the gains of the reference's experiment scripts.
"""
import os
import types

import numpy as np

from . import URDF_DIR

PANDA_URDF = os.path.join(URDF_DIR, "panda.urdf")
PANDA_WO_TOOL_URDF = os.path.join(URDF_DIR, "panda_wo_tool.urdf")
TWO_JOINT_URDF = os.path.join(URDF_DIR, "two_joint_robot.urdf")

# synthetic test robot: general revolute axes, x/y/z prismatic joints, multi-axis rpy, three branchings (urdf/make_urdf.py);
# its joint order is scrambled on purpose (kinematics.py:197: q is re-ordered per frame)
GANTRY_URDF = os.path.join(URDF_DIR, "gantry_arm.urdf")
GANTRY_ORDER = ["wrist", "slide_x", "aux_tip_joint", "elbow", "finger_z", "shoulder", "probe_slide", "aux_arm", "wrist_roll"]
GANTRY_PRISMATIC = ("slide_x", "probe_slide", "finger_z")
GANTRY_EULER_GOAL = [0.3, -0.2, 0.5]

PANDA_ORDER_7 = [f"panda_joint{i}" for i in range(1, 8)]
PANDA_ORDER_9 = PANDA_ORDER_7 + ["panda_finger_joint1", "panda_finger_joint2"]
TWO_JOINT_ORDER = ["joint_1", "joint_2"]
EE_FRAME = "panda_grasptarget_hand"

# reference: simulation.py:137-139 (12 entries, indexed by idx_controllable = [0..6, 9, 10])
_Q_LIM_LOW_12 = np.array([-2.9671, -1.8326, -2.9671, -3.1416, -2.9671, -0.0873, -2.9671, 0.0, 0.0, 0.0, 0.0, 0.0])
_Q_LIM_HIGH_12 = np.array([2.9671, 1.8326, 2.9671, 0.0, 2.9671, 3.8223, 2.9671, 0.0, 0.0, 0.04, 0.04, 0.0])
_IDX_CONTROLLABLE = [0, 1, 2, 3, 4, 5, 6, 9, 10]
PANDA_Q_LOW = _Q_LIM_LOW_12[_IDX_CONTROLLABLE]
PANDA_Q_HIGH = _Q_LIM_HIGH_12[_IDX_CONTROLLABLE]
PANDA_Q_READY = np.array([0, -0.3, 0, -2.2, 0, 2.0, np.pi / 4, 0.02, 0.02])
# reference: experiments/franka_panda/06_cluttered_environment.py:89
CSPACE_GOAL_9 = np.array([0.0, -0.9, 0.0, -2.8, 0.0, 2.0, 0.7853981633974483, 0.02, 0.02])
# reference: experiments/franka_panda/04_nullspace_control.py:51
NULLSPACE_Q0_9 = np.array([np.pi / 2, -0.05, 0, -2.01, 0, 2.22, 0.79, 0.02, 0.02])


def product_namespace():
    from . import data_management, kinematics, rmp, rmp2, taskmap
    ns = types.SimpleNamespace()
    for mod in (kinematics, taskmap, rmp, rmp2, data_management):
        for k, v in vars(mod).items():
            if not k.startswith("_"):
                setattr(ns, k, v)
    return ns


# ---------------------------------------------------------------------------------------- leaves
def ee_position_taskmap(ns, fkine, frame=EE_FRAME):
    return ns.chain_taskmaps([ns.TaskmapByForwardKinematic(fkine, frame), ns.TaskmapFrom4x4ToPosition()])


def target_attractor(ns, fkine, goal, frame=EE_FRAME):
    """gains of experiments/franka_panda/06_cluttered_environment.py:70-75"""
    return ns.TargetAttractor(
        goal=goal, accel_p_gain=0.3, accel_d_gain=0.6, accel_norm_eps=0.075, metric_alpha_length_scale=0.05,
        min_metric_alpha=0.03, max_metric_scalar=1, min_metric_scalar=0.5, proximity_metric_boost_scalar=1.,
        proximity_metric_boost_length_scale=0.02, taskmap=ee_position_taskmap(ns, fkine, frame), name='attractor')


def obstacle_leaf(ns, taskmap, frame):
    """gains of experiments/franka_panda/06_cluttered_environment.py:109-115"""
    return ns.ObstacleAvoidance(
        margin=0., damping_gain=50, damping_std_dev=0.04, damping_robustness_eps=0.01,
        damping_velocity_gate_length_scale=0.01, repulsion_gain=800, repulsion_std_dev=0.01,
        metric_modulation_radius=0.5, metric_scalar=1, metric_exploder_std_dev=0.02, metric_exploder_eps=0.001,
        taskmap=taskmap, name=f'collision_avoidance_for_{frame}')


def collision_frames(fkine):
    return [name for name, has in zip(fkine.frame_names, fkine.has_collision) if has]


def add_obstacle_leaves(ns, core, fkine, distance_taskmap_for):
    """One ObstacleAvoidance leaf per collision frame (06_cluttered_environment.py:95-116).
    ``distance_taskmap_for(frame)`` returns the second stage of the chain."""
    for frame in collision_frames(fkine):
        tm = ns.chain_taskmaps([ns.TaskmapByForwardKinematic(fkine, frame), distance_taskmap_for(frame)])
        core.add_rmp(obstacle_leaf(ns, tm, frame))


# ----------------------------------------------------------------------------------------- trees
def build_config1(ns, fkine, goal):
    """two-joint target reaching (experiments/two_joint_robot/01_target_rmp_only.py:40-46)"""
    core = ns.RmpCore()
    tm = ee_position_taskmap(ns, fkine, 'link_23')
    core.add_rmp(ns.TargetPolicy(alpha=0.1, beta=0.5, c=0.1, goal=goal, name='target', taskmap=tm))
    return core


def build_config2(ns, fkine, goal, n):
    """Panda target + joint-space biasing (experiments/franka_panda/04_nullspace_control.py:43-52)"""
    core = ns.RmpCore()
    core.add_rmp(ns.TargetPolicy(alpha=0.1, beta=1, c=0.1, goal=goal, name='target',
                                 taskmap=ee_position_taskmap(ns, fkine)))
    core.add_rmp(ns.ConfigurationSpaceBiasing(gamma_p=0.01, gamma_d=0.1, q0=NULLSPACE_Q0_9[:n],
                                              name='jointspace_biasing', w=0.05))
    return core


def build_config3(ns, fkine, goal, n, distance_taskmap_for):
    """cluttered environment tree (experiments/franka_panda/06_cluttered_environment.py:61-116)"""
    core = ns.RmpCore()
    core.add_rmp(target_attractor(ns, fkine, goal))
    core.add_rmp(ns.JointVelocityCap(max_velocity=0.5, velocity_damping_region=0.15, damping_gain=5.0,
                                     metric_weight=0.05))
    core.add_rmp(ns.JointDamping(accel_d_gain=1, metric_scalar=0.005, inertia=0.3))
    core.add_rmp(ns.CSpaceBiasing(goal=CSPACE_GOAL_9[:n], metric_scalar=0.005, position_gain=1, damping_gain=2,
                                  robust_position_term_thresh=0.5, inertia=0.0001))
    add_obstacle_leaves(ns, core, fkine, distance_taskmap_for)
    return core


def build_config4(ns, fkine, goal, n, distance_taskmap_for):
    """target + joint limits + obstacles (BASELINE.json configs[3]; limit gains of
    experiments/two_joint_robot/03_jointlimit_avoiding.py:36)"""
    core = ns.RmpCore()
    core.add_rmp(target_attractor(ns, fkine, goal))
    core.add_rmp(ns.JointLimitAvoidance(PANDA_Q_LOW[:n], PANDA_Q_HIGH[:n], gamma_p=0.3, gamma_d=1))
    add_obstacle_leaves(ns, core, fkine, distance_taskmap_for)
    return core


def build_config5(ns, fkine, goal, n, distance_taskmap_for):
    """full tree = config 3 leaves + joint limits (BASELINE.json configs[4])"""
    core = build_config3(ns, fkine, goal, n, distance_taskmap_for)
    core.add_rmp(ns.JointLimitAvoidance(PANDA_Q_LOW[:n], PANDA_Q_HIGH[:n], gamma_p=0.3, gamma_d=1))
    return core


def euler_taskmap(ns, fkine, frame):
    return ns.chain_taskmaps([ns.TaskmapByForwardKinematic(fkine, frame), ns.TaskmapFrom4x4ToEuler()])


def build_config6(ns, fkine, goal, n, distance_taskmap_for):
    """Test-only tree on the synthetic gantry arm: position + ORIENTATION (Euler task map, taskmap.py:57-67)
    targets on the tool tip, a second position target on the other branch, damping, obstacle leaves on every
    collision frame."""
    core = ns.RmpCore()
    core.add_rmp(ns.TargetPolicy(alpha=0.1, beta=0.5, c=0.1, goal=goal, name='target',
                                 taskmap=ee_position_taskmap(ns, fkine, 'tool')))
    core.add_rmp(ns.TargetPolicy(alpha=0.2, beta=0.4, c=0.1, goal=GANTRY_EULER_GOAL, name='orientation',
                                 taskmap=euler_taskmap(ns, fkine, 'tool')))
    core.add_rmp(target_attractor(ns, fkine, [0.2, 0.3, 0.4], frame='aux_tip_joint'))
    core.add_rmp(ns.JointDamping(accel_d_gain=1, metric_scalar=0.005, inertia=0.05))
    add_obstacle_leaves(ns, core, fkine, distance_taskmap_for)
    return core


BUILDERS = {2: build_config2, 3: build_config3, 4: build_config4, 5: build_config5, 6: build_config6}
N_SPHERES = {2: 0, 3: 16, 4: 64, 5: 64, 6: 8}
SEEDS = {1: 0, 2: 1, 3: 2, 4: 3, 5: 4, 6: 5}
FULL_BATCH = {1: 1, 2: 4096, 3: 65536, 4: 1 << 20, 5: 8 << 20}


# ---------------------------------------------------------------------------------------- inputs
def sample_two_joint(B, seed=0):
    """SURVEY.md section 8d config 1"""
    rng = np.random.RandomState(seed)
    q = rng.uniform(-np.pi, np.pi, size=(B, 2)).astype(np.float32)
    qd = rng.uniform(-0.5, 0.5, size=(B, 2)).astype(np.float32)
    goal = rng.uniform([0.1, -1.4, 0.1], [1.4, 1.4, 0.1], size=(B, 3)).astype(np.float32)
    return q, qd, goal


def sample_panda_state(B, n, seed):
    """q ~ U(limits), qd ~ U(-0.3, 0.3), goal ~ U([0.3,-0.7,0.3],[0.7,0.7,0.7])
    (goal box: experiments/franka_panda/01_target_rmp_only.py:60)"""
    rng = np.random.RandomState(seed)
    q = rng.uniform(PANDA_Q_LOW[:n], PANDA_Q_HIGH[:n], size=(B, n)).astype(np.float32)
    qd = rng.uniform(-0.3, 0.3, size=(B, n)).astype(np.float32)
    goal = rng.uniform([0.3, -0.7, 0.3], [0.7, 0.7, 0.7], size=(B, 3)).astype(np.float32)
    return q, qd, goal


def sample_gantry_state(B, seed):
    """revolute joints ~ U(-2.5, 2.5), prismatic ~ U(-0.2, 0.3) in GANTRY_ORDER, qd ~ U(-0.3, 0.3), goal in reach"""
    rng = np.random.RandomState(seed)
    lo = np.array([-0.2 if name in GANTRY_PRISMATIC else -2.5 for name in GANTRY_ORDER])
    hi = np.array([0.3 if name in GANTRY_PRISMATIC else 2.5 for name in GANTRY_ORDER])
    q = rng.uniform(lo, hi, size=(B, len(GANTRY_ORDER))).astype(np.float32)
    qd = rng.uniform(-0.3, 0.3, size=(B, len(GANTRY_ORDER))).astype(np.float32)
    goal = rng.uniform([-0.4, -0.4, 0.2], [0.6, 0.4, 0.8], size=(B, 3)).astype(np.float32)
    return q, qd, goal


def sample_spheres(B, O, seed, frame_origins=None, min_gap=0.03, rounds=30):
    """O spheres per environment: centre ~ U([-0.8,-0.8,0],[0.8,0.8,1.2]), radius ~ U(0.025, 0.1);
    a sphere closer than ``min_gap`` (surface distance) to any of ``frame_origins`` [B,K,3] is redrawn."""
    rng = np.random.RandomState(seed + 1000)
    lo, hi = np.array([-0.8, -0.8, 0.0]), np.array([0.8, 0.8, 1.2])
    sph = np.concatenate([rng.uniform(lo, hi, size=(B, O, 3)), rng.uniform(0.025, 0.1, size=(B, O, 1))], axis=-1)
    if frame_origins is not None and O > 0:
        P = np.asarray(frame_origins, dtype=np.float64)
        for _ in range(rounds):
            d = np.linalg.norm(sph[:, :, None, :3] - P[:, None, :, :], axis=-1) - sph[:, :, None, 3]
            bad = (d < min_gap).any(-1)
            nbad = int(bad.sum())
            if nbad == 0:
                break
            sph[bad] = np.concatenate([rng.uniform(lo, hi, size=(nbad, 3)), rng.uniform(0.025, 0.1, size=(nbad, 1))], -1)
        else:
            sph[bad] = np.array([3.0, 3.0, 3.0, 0.05])       # park the stragglers out of reach
    return sph.astype(np.float32)


def closest_points_on_spheres(origins, spheres):
    """Pairs the reference's distance feed would report for sphere obstacles and frame-origin
    control points: pos_on_link = origin [K,3] (broadcast over O), pos_on_obstacle = centre +
    radius * unit(origin - centre).  origins [K,3], spheres [O,4] -> ([K,O,3], [K,O,3])."""
    origins = np.asarray(origins, dtype=np.float32)
    spheres = np.asarray(spheres, dtype=np.float32)
    r = origins[:, None, :] - spheres[None, :, :3]
    dist = np.linalg.norm(r, axis=-1, keepdims=True)
    on_obst = spheres[None, :, :3] + spheres[None, :, 3:4] * r / dist
    on_link = np.broadcast_to(origins[:, None, :], on_obst.shape).copy()
    return on_link.astype(np.float32), on_obst.astype(np.float32)


def synth_inputs_device(fk, n, B, O_, n_buffers, seed, device):
    """Seeded synthetic environments generated on the device (distribution: SURVEY.md section 8d)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lo = torch.tensor(PANDA_Q_LOW[:n], dtype=torch.float32, device=device)
    hi = torch.tensor(PANDA_Q_HIGH[:n], dtype=torch.float32, device=device)
    q = lo + (hi - lo) * torch.rand(B, n, generator=g, device=device)
    qd = -0.3 + 0.6 * torch.rand(B, n, generator=g, device=device)
    glo = torch.tensor([0.3, -0.7, 0.3], device=device)
    ghi = torch.tensor([0.7, 0.7, 0.7], device=device)
    goal = glo + (ghi - glo) * torch.rand(B, 3, generator=g, device=device)
    spheres = []
    if O_:
        frames = collision_frames(fk)
        origins = torch.stack([fk.forward(q, fr)[:, :3, 3] for fr in frames], dim=1)       # [B,K,3] CUDA FK kernel
        slo = torch.tensor([-0.8, -0.8, 0.0], device=device)
        shi = torch.tensor([0.8, 0.8, 1.2], device=device)

        def draw(count):
            c = slo + (shi - slo) * torch.rand(count, 3, generator=g, device=device)
            r = 0.025 + 0.075 * torch.rand(count, 1, generator=g, device=device)
            return torch.cat([c, r], dim=-1)

        for _ in range(n_buffers):
            sph = draw(B * O_).reshape(B, O_, 4)
            for _round in range(30):
                bad = torch.zeros(B, O_, dtype=torch.bool, device=device)
                for k in range(origins.shape[1]):                                         # keeps the temporaries [B,O]
                    d = torch.linalg.norm(sph[:, :, :3] - origins[:, None, k, :], dim=-1) - sph[:, :, 3]
                    bad |= d < 0.03
                nbad = int(bad.sum())
                if nbad == 0:
                    break
                sph[bad] = draw(nbad)
            spheres.append(sph.contiguous())
    return q.contiguous(), qd.contiguous(), goal.contiguous(), spheres


def closed_loop_scene_device(fk, n, B, O_, seed, device):
    """Scenes a closed loop is well behaved in (the situation of experiments/franka_panda/06_cluttered_environment.py),
    generated on the device: start near the ready pose at rest, a goal inside the workspace, O_ spheres anywhere around
    but at least 0.12 m (surface) from every collision frame of the start pose and from the goal.
    -> q0 [B,n], qd0 [B,n], goal [B,3], spheres [B,O_,4]."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rand = lambda *shape: torch.rand(*shape, generator=g, device=device)
    q0 = torch.tensor(PANDA_Q_READY[:n], dtype=torch.float32, device=device) + (-0.05 + 0.1 * rand(B, n))
    glo, ghi = torch.tensor([0.3, -0.35, 0.25], device=device), torch.tensor([0.6, 0.35, 0.65], device=device)
    goal = glo + (ghi - glo) * rand(B, 3)
    origins = torch.stack([fk.forward(q0, fr)[:, :3, 3] for fr in collision_frames(fk)], dim=1)      # [B,K,3]
    slo, shi = torch.tensor([-0.2, -0.6, 0.0], device=device), torch.tensor([0.8, 0.6, 1.0], device=device)

    def draw(count):
        return torch.cat([slo + (shi - slo) * rand(count, 3), 0.03 + 0.05 * rand(count, 1)], dim=-1)

    sph = draw(B * O_).reshape(B, O_, 4)
    for _round in range(40):
        bad = torch.linalg.norm(sph[:, :, :3] - goal[:, None, :], dim=-1) - sph[:, :, 3] < 0.12
        for k in range(origins.shape[1]):
            bad |= torch.linalg.norm(sph[:, :, :3] - origins[:, None, k, :], dim=-1) - sph[:, :, 3] < 0.12
        nbad = int(bad.sum())
        if nbad == 0:
            break
        sph[bad] = draw(nbad)
    else:
        sph[bad] = torch.tensor([3.0, 3.0, 3.0, 0.05], device=device)
    return q0.contiguous(), torch.zeros_like(q0), goal.contiguous(), sph.contiguous()
