"""RMP2-style leaf policies -- host-side mirror of the reference's ``rmp2.py``.

Constructor signatures and attribute names are the reference's; the arithmetic lives in
csrc/rmp2_leaves.cuh.
"""
from . import _native
from ._leaf import RiemannianMotionPolicy, as_float_list
from .taskmap import IdentityTaskmap


class TargetAttractor(RiemannianMotionPolicy):
    """reference: rmp2.py:31-83."""
    leaf_type = _native.LEAF_TARGET_ATTRACTOR

    def __init__(self, goal, accel_p_gain, accel_d_gain, accel_norm_eps, metric_alpha_length_scale,
                 min_metric_alpha, max_metric_scalar, min_metric_scalar, proximity_metric_boost_scalar,
                 proximity_metric_boost_length_scale, taskmap, name='attractor'):
        super().__init__(name, taskmap)
        self.goal = goal
        self.accel_p_gain = accel_p_gain
        self.accel_d_gain = accel_d_gain
        self.accel_norm_eps = accel_norm_eps
        self.metric_alpha_length_scale = metric_alpha_length_scale
        self.min_metric_alpha = min_metric_alpha
        self.max_metric_scalar = max_metric_scalar
        self.min_metric_scalar = min_metric_scalar
        self.proximity_metric_boost_scalar = proximity_metric_boost_scalar
        self.proximity_metric_boost_length_scale = proximity_metric_boost_length_scale

    def _params(self):
        return [self.accel_p_gain, self.accel_d_gain, self.accel_norm_eps, self.metric_alpha_length_scale,
                self.min_metric_alpha, self.max_metric_scalar, self.min_metric_scalar,
                self.proximity_metric_boost_scalar, self.proximity_metric_boost_length_scale]

    def _vec(self, dim):
        return as_float_list(self.goal, 3, "TargetAttractor.goal")


class JointVelocityCap(RiemannianMotionPolicy):
    """reference: rmp2.py:86-112."""
    leaf_type = _native.LEAF_VELOCITY_CAP

    def __init__(self, max_velocity, velocity_damping_region, damping_gain, metric_weight,
                 name='joint_velocity_cap'):
        super().__init__(name, taskmap=IdentityTaskmap())
        self.max_velocity = max_velocity
        self.velocity_damping_region = velocity_damping_region
        self.damping_gain = damping_gain
        self.metric_weight = metric_weight
        self.eps = 1e-6
        self.damped_velocity_cutoff = self.max_velocity - self.velocity_damping_region

    def _params(self):
        return [self.max_velocity, self.velocity_damping_region, self.damping_gain, self.metric_weight]


class JointDamping(RiemannianMotionPolicy):
    """reference: rmp2.py:115-137."""
    leaf_type = _native.LEAF_JOINT_DAMPING

    def __init__(self, accel_d_gain, metric_scalar, inertia, name='joint_damping'):
        super().__init__(name=name, taskmap=IdentityTaskmap())
        self.accel_d_gain = accel_d_gain
        self.metric_scalar = metric_scalar
        self.inertia = inertia

    def _params(self):
        return [self.accel_d_gain, self.metric_scalar, self.inertia]


class ObstacleAvoidance(RiemannianMotionPolicy):
    """reference: rmp2.py:140-196."""
    leaf_type = _native.LEAF_OBSTACLE_AVOIDANCE

    def __init__(self, margin, damping_gain, damping_std_dev, damping_robustness_eps,
                 damping_velocity_gate_length_scale, repulsion_gain, repulsion_std_dev,
                 metric_modulation_radius, metric_scalar, metric_exploder_std_dev, metric_exploder_eps,
                 taskmap, name):
        super().__init__(name=name, taskmap=taskmap)
        self.margin = margin
        self.damping_gain = damping_gain
        self.damping_std_dev = damping_std_dev
        self.damping_robustness_eps = damping_robustness_eps
        self.damping_velocity_gate_length_scale = damping_velocity_gate_length_scale
        self.repulsion_gain = repulsion_gain
        self.repulsion_std_dev = repulsion_std_dev
        self.metric_modulation_radius = metric_modulation_radius
        self.metric_scalar = metric_scalar
        self.metric_exploder_std_dev = metric_exploder_std_dev
        self.metric_exploder_eps = metric_exploder_eps

    def _params(self):
        return [self.margin, self.damping_gain, self.damping_std_dev, self.damping_robustness_eps,
                self.damping_velocity_gate_length_scale, self.repulsion_gain, self.repulsion_std_dev,
                self.metric_modulation_radius, self.metric_scalar, self.metric_exploder_std_dev,
                self.metric_exploder_eps]


class CSpaceBiasing(RiemannianMotionPolicy):
    """Configuration-space target reaching (reference: rmp2.py:198-226)."""
    leaf_type = _native.LEAF_CSPACE_BIASING

    def __init__(self, goal, metric_scalar, position_gain, damping_gain, robust_position_term_thresh,
                 inertia, taskmap=None, name='cspace_target'):
        super().__init__(name=name, taskmap=IdentityTaskmap() if taskmap is None else taskmap)
        self.goal = goal
        self.metric_scalar = metric_scalar
        self.position_gain = position_gain
        self.damping_gain = damping_gain
        self.robust_position_term_thresh = robust_position_term_thresh
        self.inertia = inertia

    def _params(self):
        return [self.metric_scalar, self.position_gain, self.damping_gain, self.robust_position_term_thresh,
                self.inertia]

    def _vec(self, dim):
        return as_float_list(self.goal, dim, "CSpaceBiasing.goal")
