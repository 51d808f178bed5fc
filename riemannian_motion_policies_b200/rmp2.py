"""RMP2-style leaf policies -- host-side mirror of the reference's ``rmp2.py``.

Constructor signatures (argument names, order, defaults) and attribute names are the reference's; they are
declared as data (``FIELDS``) and bound by ``_leaf.DeclaredLeaf``.  The arithmetic lives in
csrc/rmp2_leaves.cuh; ``PARAMS`` is the order the C ABI documents for each leaf (include/rmp2_b200.h).
"""
from . import _native
from ._leaf import DeclaredLeaf, as_float_list


class TargetAttractor(DeclaredLeaf):
    """reference: rmp2.py:31-83."""
    leaf_type = _native.LEAF_TARGET_ATTRACTOR
    PARAMS = ("accel_p_gain", "accel_d_gain", "accel_norm_eps", "metric_alpha_length_scale", "min_metric_alpha",
              "max_metric_scalar", "min_metric_scalar", "proximity_metric_boost_scalar",
              "proximity_metric_boost_length_scale")
    FIELDS = ("goal",) + PARAMS + ("taskmap", ("name", "attractor"))

    def _vec(self, dim):
        return as_float_list(self.goal, 3, "TargetAttractor.goal")


class JointVelocityCap(DeclaredLeaf):
    """reference: rmp2.py:86-112."""
    leaf_type = _native.LEAF_VELOCITY_CAP
    PARAMS = ("max_velocity", "velocity_damping_region", "damping_gain", "metric_weight")
    FIELDS = PARAMS + (("name", "joint_velocity_cap"),)

    def _post_init(self):
        self.eps = 1e-6                                                        # rmp2.py:96-97
        self.damped_velocity_cutoff = self.max_velocity - self.velocity_damping_region


class JointDamping(DeclaredLeaf):
    """reference: rmp2.py:115-137."""
    leaf_type = _native.LEAF_JOINT_DAMPING
    PARAMS = ("accel_d_gain", "metric_scalar", "inertia")
    FIELDS = PARAMS + (("name", "joint_damping"),)


class ObstacleAvoidance(DeclaredLeaf):
    """reference: rmp2.py:140-196."""
    leaf_type = _native.LEAF_OBSTACLE_AVOIDANCE
    PARAMS = ("margin", "damping_gain", "damping_std_dev", "damping_robustness_eps",
              "damping_velocity_gate_length_scale", "repulsion_gain", "repulsion_std_dev",
              "metric_modulation_radius", "metric_scalar", "metric_exploder_std_dev", "metric_exploder_eps")
    FIELDS = PARAMS + ("taskmap", "name")


class CSpaceBiasing(DeclaredLeaf):
    """Configuration-space target reaching (reference: rmp2.py:198-226)."""
    leaf_type = _native.LEAF_CSPACE_BIASING
    PARAMS = ("metric_scalar", "position_gain", "damping_gain", "robust_position_term_thresh", "inertia")
    FIELDS = ("goal",) + PARAMS + (("taskmap", None), ("name", "cspace_target"))

    def _vec(self, dim):
        return as_float_list(self.goal, dim, "CSpaceBiasing.goal")
