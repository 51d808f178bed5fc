"""ctypes binding of librmp2_b200.so (C ABI declared in include/rmp2_b200.h).

PyTorch is used only for device memory and streams: tensors are passed to the library as raw
device pointers.  There is no CPU fallback -- if the shared library is missing, or no CUDA device
is present when a compute call is made, the call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RMP2_B200_LIB: load another build of the same library (tools/ A/B studies of compile-time knobs)
LIB_PATH = os.environ.get("RMP2_B200_LIB") or os.path.join(_HERE, "csrc", "librmp2_b200.so")

RMP2_MAX_FRAMES = 24
RMP2_MAX_JOINTS = 12
RMP2_MAX_LEAVES = 40
RMP2_LEAF_PARAMS = 16
RMP2_MAX_GOAL_SLOTS = 4
RMP2_MAX_PAIR_SETS = 24

JOINT_FIXED, JOINT_REVOLUTE, JOINT_PRISMATIC = 0, 1, 2

LEAF_TARGET_POLICY = 1
LEAF_CONFIG_BIASING = 2
LEAF_JOINT_LIMIT = 3
LEAF_TARGET_ATTRACTOR = 4
LEAF_VELOCITY_CAP = 5
LEAF_JOINT_DAMPING = 6
LEAF_OBSTACLE_AVOIDANCE = 7
LEAF_CSPACE_BIASING = 8
LEAF_COLLISION_AVOIDANCE = 9

SPACE_CONFIG = 0
SPACE_FRAME_POSITION = 1
SPACE_FRAME_DISTANCE_SPHERES = 2
SPACE_FRAME_DISTANCE_PAIRS = 3
SPACE_FRAME_POINTS = 4
SPACE_FRAME_EULER = 5
PAIR_FLOATS = 8

OPT_EARLY_OUT = 0
OPT_TMA = 1
OPT_SPLIT_RESOLVE = 2
OPT_BLOCK_THREADS = 3
OPT_CHUNK_ENVS = 4
OPT_MERGE_COINCIDENT = 5
PROFILE_KERNELS = 5
SPECIALIZE_COMPILE_ONLY = 1

# every symbol include/rmp2_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "rmp2_robot_create", "rmp2_robot_destroy", "rmp2_tree_create", "rmp2_tree_destroy",
    "rmp2_tree_update_leaf", "rmp2_step", "rmp2_step_host", "rmp2_rollout", "rmp2_fk",
    "rmp2_leaf_evaluate", "rmp2_obstacle_feed", "rmp2_last_error", "rmp2_version", "rmp2_launch_count",
    "rmp2_tree_kernel_info", "rmp2_tree_profile", "rmp2_tree_profile_read", "rmp2_tree_set_option",
    "rmp2_tree_specialize", "rmp2_tree_is_specialized", "rmp2_tree_reserve", "rmp2_pinv_solve",
    "rmp2_tree_obstacle_slots",
]


class LeafDesc(ctypes.Structure):
    _fields_ = [
        ("type", ctypes.c_int32),
        ("space", ctypes.c_int32),
        ("frame", ctypes.c_int32),
        ("goal_slot", ctypes.c_int32),
        ("params", ctypes.c_float * RMP2_LEAF_PARAMS),
        ("vec", ctypes.c_float * (2 * RMP2_MAX_JOINTS)),
    ]


class StepIO(ctypes.Structure):
    _fields_ = [
        ("B", ctypes.c_int64),
        ("q", ctypes.c_void_p),
        ("qd", ctypes.c_void_p),
        ("qdd", ctypes.c_void_p),
        ("goals", ctypes.c_void_p),
        ("n_goal_slots", ctypes.c_int32),
        ("n_spheres", ctypes.c_int32),
        ("spheres", ctypes.c_void_p),
        ("pairs", ctypes.c_void_p),
        ("n_pair_sets", ctypes.c_int32),
        ("pair_counts", ctypes.c_int32 * RMP2_MAX_PAIR_SETS),
    ]


class NativeLibraryMissing(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library once.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  riemannian_motion_policies_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
    L.rmp2_robot_create.argtypes = [vp, vp, vp, vp, vp, i32, i32, ctypes.POINTER(vp)]
    L.rmp2_robot_create.restype = ctypes.c_int
    L.rmp2_robot_destroy.argtypes = [vp]
    L.rmp2_robot_destroy.restype = None
    L.rmp2_tree_create.argtypes = [vp, ctypes.POINTER(LeafDesc), i32, ctypes.POINTER(vp)]
    L.rmp2_tree_create.restype = ctypes.c_int
    L.rmp2_tree_destroy.argtypes = [vp]
    L.rmp2_tree_destroy.restype = None
    L.rmp2_tree_update_leaf.argtypes = [vp, i32, ctypes.POINTER(LeafDesc)]
    L.rmp2_tree_update_leaf.restype = ctypes.c_int
    L.rmp2_step.argtypes = [vp, ctypes.POINTER(StepIO), vp]
    L.rmp2_step.restype = ctypes.c_int
    L.rmp2_step_host.argtypes = [vp, ctypes.POINTER(StepIO)]
    L.rmp2_step_host.restype = ctypes.c_int
    L.rmp2_rollout.argtypes = [vp, ctypes.POINTER(StepIO), vp, vp, f32, i32, i32, vp]
    L.rmp2_rollout.restype = ctypes.c_int
    L.rmp2_fk.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, vp, vp]
    L.rmp2_fk.restype = ctypes.c_int
    L.rmp2_leaf_evaluate.argtypes = [ctypes.POINTER(LeafDesc), i32, i64, vp, vp, vp, vp, vp, vp]
    L.rmp2_leaf_evaluate.restype = ctypes.c_int
    L.rmp2_obstacle_feed.argtypes = [vp, vp, vp, i32, i64, vp, vp, i32, vp, i32, vp, vp, vp]
    L.rmp2_obstacle_feed.restype = ctypes.c_int
    L.rmp2_last_error.argtypes = []
    L.rmp2_last_error.restype = ctypes.c_char_p
    L.rmp2_version.argtypes = []
    L.rmp2_version.restype = ctypes.c_char_p
    L.rmp2_launch_count.argtypes = []
    L.rmp2_launch_count.restype = i64
    L.rmp2_tree_kernel_info.argtypes = [vp, i32, i32, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32),
                                        ctypes.POINTER(i32)]
    L.rmp2_tree_kernel_info.restype = ctypes.c_int
    L.rmp2_tree_set_option.argtypes = [vp, i32, i32]
    L.rmp2_tree_set_option.restype = ctypes.c_int
    L.rmp2_tree_obstacle_slots.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.rmp2_tree_obstacle_slots.restype = ctypes.c_int
    L.rmp2_tree_specialize.argtypes = [vp, i32]
    L.rmp2_tree_specialize.restype = ctypes.c_int
    L.rmp2_tree_is_specialized.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.rmp2_tree_is_specialized.restype = ctypes.c_int
    L.rmp2_tree_reserve.argtypes = [vp, i64, i32, vp]
    L.rmp2_tree_reserve.restype = ctypes.c_int
    L.rmp2_pinv_solve.argtypes = [i32, i64, vp, vp, vp, i32, i32, vp]
    L.rmp2_pinv_solve.restype = ctypes.c_int
    L.rmp2_tree_profile.argtypes = [vp, i32]
    L.rmp2_tree_profile.restype = ctypes.c_int
    L.rmp2_tree_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)]
    L.rmp2_tree_profile_read.restype = ctypes.c_int
    _lib = L
    return L


def check(rc):
    """Turn a C status into the exception the reference's callers would see."""
    if rc == 0:
        return
    msg = lib().rmp2_last_error().decode("utf-8", "replace")
    if rc == 1:
        raise ValueError(msg)
    if rc == 3:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def launch_count():
    return int(lib().rmp2_launch_count())
