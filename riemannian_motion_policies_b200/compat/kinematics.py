"""Drop-in shim: lets the reference's experiment scripts keep ``from kinematics import ...``.
Put this directory on sys.path ahead of the reference checkout."""
from riemannian_motion_policies_b200.kinematics import *  # noqa: F401,F403
