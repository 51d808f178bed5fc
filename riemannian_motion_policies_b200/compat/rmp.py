"""Drop-in shim: lets the reference's experiment scripts keep ``from rmp import ...``.
Put this directory on sys.path ahead of the reference checkout."""
from riemannian_motion_policies_b200.rmp import *  # noqa: F401,F403
from riemannian_motion_policies_b200.rmp import RiemannianMotionPolicy  # noqa: F401
