"""Executes the tree-building blocks of the REFERENCE'S experiment scripts verbatim (test infrastructure).

The north star says the two_joint_robot and franka_panda experiments "drop in unchanged".  The scripts live in
the reference checkout (never copied into this repo), so this harness only works where /root/reference exists
(the build container).  It reads a script, execs

  * its import header (everything before the first ``def``), unchanged,
  * the one line that creates ``goal`` (when the script has one), unchanged,
  * the block from ``# forward kinematic`` up to ``# simulation`` -- kinematics, Datamanager, task maps, leaves,
    ``core.add_rmp`` -- unchanged,

in a namespace where only the *simulation side* is stubbed: ``pybullet`` (joint table and collision flags served
from the URDF; PyBullet's joint order equals the URDF order), ``simulation`` (plain Goal /
Cylinder / robot value objects), ``tensorflow`` (oracle/tf_shim: the scripts call ``tf.constant``).  Which
implementation the block builds on is decided exactly as INTEGRATION.md section A says -- by what comes first on
``sys.path``:  ``riemannian_motion_policies_b200/compat`` (the product) or the reference checkout itself (the
reference's own modules under the shim, used to generate tests/golden/ref_exp_*.npz).
"""
import contextlib
import os
import sys
import textwrap
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
COMPAT = os.path.join(ROOT, "riemannian_motion_policies_b200", "compat")
TF_SHIM = os.path.join(ROOT, "oracle", "tf_shim")

PANDA_LIMITS = (np.array([-2.9671, -1.8326, -2.9671, -3.1416, -2.9671, -0.0873, -2.9671, 0.0, 0.0, 0.0, 0.0, 0.0]),
                np.array([2.9671, 1.8326, 2.9671, 0.0, 2.9671, 3.8223, 2.9671, 0.0, 0.0, 0.04, 0.04, 0.0]))

# script (relative to the reference checkout) -> robot the script's ``robot`` object stands for
SCRIPTS = {
    "exp06": ("experiments/franka_panda/06_cluttered_environment.py", "panda"),
    "exp04": ("experiments/franka_panda/04_nullspace_control.py", "panda"),
    "two01": ("experiments/two_joint_robot/01_target_rmp_only.py", "two_joint"),
    "two03": ("experiments/two_joint_robot/03_jointlimit_avoiding.py", "two_joint"),
    "two05": ("experiments/two_joint_robot/05_obstacle_avoidance.py", "two_joint"),
}
URDF_OF = {"panda": "urdf/franka_panda/panda.urdf", "two_joint": "urdf/TwoJointRobot_wo_fixedJoints.urdf"}
MODULES = ("rmp", "rmp2", "kinematics", "taskmap", "data_management", "simulation", "pybullet", "pybullet_data",
           "imageio", "tensorflow", "helper", "helper.pybullet_helper", "helper.rmp_helper", "helper.tensorflow_helper",
           "helper.urdf_parsing", "helper.trigonometry_helper", "experiments")


def available():
    return os.path.isdir(REFERENCE)


def _joint_table(robot):
    """(name, movable, child link has collision) per URDF joint, file order."""
    from xml.etree import ElementTree
    root = ElementTree.parse(os.path.join(REFERENCE, URDF_OF[robot])).getroot()
    links = {ln.attrib["name"]: ln for ln in root.findall("link")}
    out = []
    for j in root.findall("joint"):
        col = links[j.find("child").attrib["link"]].find("collision")
        out.append((j.attrib["name"], j.attrib["type"] != "fixed", col is not None and len(col) > 0))
    return out


def _pybullet_stub(robot):
    table = _joint_table(robot)
    p = types.ModuleType("pybullet")
    p.getNumJoints = lambda body: len(table)

    def get_joint_info(body, i):
        name, movable, _ = table[i]
        info = [i, name.encode("ascii"), 0, (7 + i) if movable else -1] + [0] * 12 + [i - 1]
        return tuple(info)

    p.getJointInfo = get_joint_info
    p.getCollisionShapeData = lambda body, linkIndex: ((body, linkIndex, 3),) if table[linkIndex][2] else ()
    return p


def _simulation_stub(robot):
    sim = types.ModuleType("simulation")

    class _Obj:
        def __init__(self, *args, **kwargs):
            self.__dict__.update(kwargs)
            self.id = 0

    class Goal(_Obj):
        pass

    class Cylinder(_Obj):
        pass

    class Simulation(_Obj):
        pass

    class FrankaPanda(_Obj):
        q_lim_low, q_lim_high = PANDA_LIMITS
        idx_controllable = [0, 1, 2, 3, 4, 5, 6, 9, 10]

    class TwoJointRobot(_Obj):
        q_lim_low, q_lim_high = np.array([-np.pi, -np.pi]), np.array([+np.pi, +np.pi])
        idx_controllable = [0, 1]

    for cls in (Goal, Cylinder, Simulation, FrankaPanda, TwoJointRobot):
        setattr(sim, cls.__name__, cls)
    return sim, (FrankaPanda if robot == "panda" else TwoJointRobot)


@contextlib.contextmanager
def _import_context(first_on_path, robot):
    """sys.path = [first_on_path, tf shim, ..., reference checkout]; the simulation side stubbed; restored afterwards."""
    saved_path, saved_modules = list(sys.path), {k: sys.modules.get(k) for k in MODULES}
    for k in list(sys.modules):
        if k in MODULES or k.startswith(("helper.", "experiments.")):
            saved_modules.setdefault(k, sys.modules[k])
            del sys.modules[k]
    sim, robot_cls = _simulation_stub(robot)
    sys.modules["simulation"] = sim
    sys.modules["pybullet"] = _pybullet_stub(robot)
    sys.path[:] = [first_on_path, TF_SHIM] + [p for p in saved_path if p not in (first_on_path, TF_SHIM, REFERENCE)] + [REFERENCE]
    try:
        yield robot_cls
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k in MODULES or k.startswith(("helper.", "experiments.")):
                del sys.modules[k]
        for k, v in saved_modules.items():
            if v is not None:
                sys.modules[k] = v


def _slices(path):
    with open(path) as fh:
        lines = fh.read().splitlines()
    first_def = next(i for i, ln in enumerate(lines) if ln.startswith("def "))
    start = next(i for i, ln in enumerate(lines) if ln.strip().startswith("# forward kinematic"))
    stop = next(i for i, ln in enumerate(lines) if i > start and ln.strip().startswith("# simulation"))
    goal = [ln for ln in lines[first_def:start] if ln.strip().startswith("goal = Goal(")]
    header = "\n".join(ln for ln in lines[:first_def] if "sys.path.append" not in ln)
    return header, textwrap.dedent("\n".join(goal)), textwrap.dedent("\n".join(lines[start:stop])), (start + 1, stop)


def build_tree(key, implementation="product"):
    """Exec the block of script `key`; -> namespace dict (``core``, ``fkine``, maybe ``data_manager``, ``goal`` ...)
    plus ``_lines`` = the 1-based line range executed.  implementation: "product" (compat/ first on sys.path) or
    "reference" (the reference checkout first: its own rmp.py / kinematics.py ... under the TensorFlow shim)."""
    rel, robot = SCRIPTS[key]
    path = os.path.join(REFERENCE, rel)
    header, goal_line, block, line_range = _slices(path)
    first = COMPAT if implementation == "product" else REFERENCE
    with _import_context(first, robot) as robot_cls:
        env = {"__file__": path, "__name__": "dropin_" + key}
        exec(compile(header, path, "exec"), env)
        owner = env["RmpCore"].__module__
        assert owner.startswith("riemannian_motion_policies_b200") == (implementation == "product"), owner
        if implementation == "reference":
            env["RmpCore"] = (lambda cls: (lambda: cls(rmps={})))(env["RmpCore"])   # the reference's default dict is shared
        env["robot"] = robot_cls()
        if goal_line:
            exec(compile(goal_line, path, "exec"), env)
        import builtins
        real_print, builtins.print = builtins.print, (lambda *a, **k: None)         # the reference prints on every FK build
        try:
            exec(compile(block, path, "exec"), env)
        finally:
            builtins.print = real_print
    env["_lines"] = line_range
    return env
