"""Empty stand-in: the reference imports pybullet at module top (kinematics.py:4) but the control-step
path never calls it.  TEST INFRASTRUCTURE ONLY."""
GUI = 1
DIRECT = 2
