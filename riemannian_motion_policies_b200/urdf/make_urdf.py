"""Emit the kinematic-only URDF files this repo ships (no meshes, no inertia).

The robots are the ones the reference experiments use
(reference: urdf/TwoJointRobot_wo_fixedJoints.urdf, urdf/franka_panda/panda.urdf,
urdf/franka_panda/panda_wo_tool.urdf).  Only what the control-step hot path reads
is kept: joint name/type/origin/axis/limits, parent/child links and whether the
child link carries a <collision> element (reference: helper/urdf_parsing.py:87-93).
The numbers are the public Franka Emika Panda kinematic constants (SURVEY.md
Appendix B); tests/test_urdf.py checks that these files parse to exactly the same
frame table as the reference's own URDFs whenever /root/reference is present.

Run:  python -m riemannian_motion_policies_b200.urdf.make_urdf
"""
import os

HALF_PI = "1.57079632679"

# (joint, type, parent link, child link, rpy, xyz, axis, (lower, upper, velocity), child has collision)
PANDA = [
    ("panda_joint1", "revolute", "panda_link0", "panda_link1", "0 0 0", "0 0 0.333", "0 0 1", (-2.9671, 2.9671, 2.1750), True),
    ("panda_joint2", "revolute", "panda_link1", "panda_link2", f"-{HALF_PI} 0 0", "0 0 0", "0 0 1", (-1.8326, 1.8326, 2.1750), True),
    ("panda_joint3", "revolute", "panda_link2", "panda_link3", f"{HALF_PI} 0 0", "0 -0.316 0", "0 0 1", (-2.9671, 2.9671, 2.1750), True),
    ("panda_joint4", "revolute", "panda_link3", "panda_link4", f"{HALF_PI} 0 0", "0.0825 0 0", "0 0 1", (-3.1416, 0.0, 2.1750), True),
    ("panda_joint5", "revolute", "panda_link4", "panda_link5", f"-{HALF_PI} 0 0", "-0.0825 0.384 0", "0 0 1", (-2.9671, 2.9671, 2.6100), True),
    ("panda_joint6", "revolute", "panda_link5", "panda_link6", f"{HALF_PI} 0 0", "0 0 0", "0 0 1", (-0.0873, 3.8223, 2.6100), True),
    ("panda_joint7", "revolute", "panda_link6", "panda_link7", f"{HALF_PI} 0 0", "0.088 0 0", "0 0 1", (-2.9671, 2.9671, 2.6100), True),
    ("panda_joint8", "fixed", "panda_link7", "panda_link8", "0 0 0", "0 0 0.107", None, None, False),
    ("panda_hand_joint", "fixed", "panda_link8", "panda_hand", "0 0 -0.785398163397", "0 0 0", None, None, True),
    ("panda_finger_joint1", "prismatic", "panda_hand", "panda_leftfinger", "0 0 0", "0 0 0.0584", "0 1 0", (0.0, 0.04, 0.2), True),
    ("panda_finger_joint2", "prismatic", "panda_hand", "panda_rightfinger", "0 0 0", "0 0 0.0584", "0 -1 0", (0.0, 0.04, 0.2), True),
    ("panda_grasptarget_hand", "fixed", "panda_hand", "panda_grasptarget", "0 0 0", "0 0 0.105", None, None, False),
]
PANDA_WO_TOOL = [j for j in PANDA if "finger" not in j[0]]

TWO_JOINT = [
    ("joint_1", "revolute", "base_link", "link_1", "0 0 0", "0 0 0.075", "0 0 1", (-3.14, 3.14, 5.0), True),
    ("joint_2", "revolute", "link_1", "link_2", "0 0 0", "1.0 0. 0.05", "0 0 1", (-3.14, 3.14, 5.0), True),
    ("link_23", "fixed", "link_2", "link_23_cyl", "0 0 0", "1.0 0 0", None, None, True),
]


# Synthetic test robot (not in the reference): everything the Panda files do not exercise -- revolute joints about
# general unit axes, prismatic joints along x / y / z, constant rotations with all three rpy components (where the
# reference's R_x R_y R_z order matters, kinematics.py:123-127), and a kinematic tree that branches three times.
GANTRY = [
    ("slide_x", "prismatic", "base", "carriage", "0.1 -0.2 0.3", "0 0 0.1", "1 0 0", (-0.2, 0.3, 1.0), True),
    ("shoulder", "revolute", "carriage", "upper", "0.3 0.4 -0.5", "0.05 0.02 0.2", "0.36 0.48 0.8", (-2.5, 2.5, 2.0), True),
    ("elbow", "revolute", "upper", "fore", "-0.7 0.2 0.9", "0.3 0 0.05", "0.6 0 0.8", (-2.5, 2.5, 2.0), True),
    ("aux_arm", "revolute", "upper", "aux_link", "0.4 -0.3 0.2", "-0.1 0.15 0.1", "-0.48 0.6 0.64", (-2.5, 2.5, 2.0), True),
    ("wrist", "revolute", "fore", "wrist_link", "0.2 -1.1 0.4", "0.25 0.05 0", "0 1 0", (-2.5, 2.5, 2.0), True),
    ("probe_slide", "prismatic", "fore", "probe", "-0.3 0.6 0.1", "0.1 -0.1 0.05", "0 1 0", (-0.2, 0.3, 1.0), True),
    ("aux_tip_joint", "revolute", "aux_link", "aux_tip", "0.9 0.1 -0.6", "0.2 0 0", "1 0 0", (-2.5, 2.5, 2.0), True),
    ("wrist_roll", "revolute", "wrist_link", "hand", "-0.2 0.3 0.7", "0 0 0.12", "0.666666667 0.666666667 0.333333333", (-2.5, 2.5, 2.0), True),
    ("finger_z", "prismatic", "hand", "finger", "0.1 0.2 -0.3", "0.02 0 0.05", "0 0 1", (-0.2, 0.3, 1.0), True),
    ("tool", "fixed", "hand", "tool_tip", "0.5 0.5 0.5", "0 0.03 0.1", None, None, False),
]


def emit(robot_name, base_link, joints, base_has_collision=True):
    """Return URDF text: all links first, then joints in chain order."""
    collision = ('    <collision>\n      <geometry>\n        <sphere radius="0.05"/>\n'
                 '      </geometry>\n    </collision>\n')
    out = ['<?xml version="1.0" ?>', f'<robot name="{robot_name}">']
    links = [(base_link, base_has_collision)] + [(j[3], j[8]) for j in joints]
    for name, has_col in links:
        out.append(f'  <link name="{name}">')
        if has_col:
            out.append(collision.rstrip("\n"))
        out.append("  </link>")
    for name, jtype, parent, child, rpy, xyz, axis, lim, _ in joints:
        out.append(f'  <joint name="{name}" type="{jtype}">')
        out.append(f'    <origin rpy="{rpy}" xyz="{xyz}"/>')
        out.append(f'    <parent link="{parent}"/>')
        out.append(f'    <child link="{child}"/>')
        if axis is not None:
            out.append(f'    <axis xyz="{axis}"/>')
        if lim is not None:
            out.append(f'    <limit effort="100" lower="{lim[0]}" upper="{lim[1]}" velocity="{lim[2]}"/>')
        out.append("  </joint>")
    out.append("</robot>")
    return "\n".join(out) + "\n"


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    files = {
        "panda.urdf": emit("panda", "panda_link0", PANDA),
        "panda_wo_tool.urdf": emit("panda", "panda_link0", PANDA_WO_TOOL),
        "two_joint_robot.urdf": emit("TwoJointRobot", "base_link", TWO_JOINT),
        "gantry_arm.urdf": emit("gantry_arm", "base", GANTRY),
    }
    for fname, text in files.items():
        with open(os.path.join(here, fname), "w") as fh:
            fh.write(text)
        print("wrote", fname)


if __name__ == "__main__":
    main()
