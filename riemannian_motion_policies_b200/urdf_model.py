"""Host-side URDF -> frame table (init time only, never on the step path).

Produces, for every joint of a URDF, one *frame* in the same order the reference
enumerates them (reference: helper/urdf_parsing.py:57-97 breadth-first from the base
link, joints visited in file order; helper/urdf_parsing.py:134-147 one backward path
per non-root element; kinematics.py:169-171 ``frame_names`` = last entry of each path).

The constant transform of a frame follows the reference's conventions, including
its roll/pitch/yaw order ``R_x(r) @ R_y(p) @ R_z(y)`` (reference: kinematics.py:123-127,
200-203), evaluated in float32.
"""
from collections import deque
from dataclasses import dataclass, field
from typing import List, Optional
from xml.etree import ElementTree

import numpy as np

JOINT_FIXED, JOINT_REVOLUTE, JOINT_PRISMATIC = 0, 1, 2
# the reference's type masks know exactly these three (kinematics.py:205-209)
_TYPE_CODE = {"fixed": JOINT_FIXED, "revolute": JOINT_REVOLUTE, "prismatic": JOINT_PRISMATIC}


@dataclass
class Frame:
    name: str
    link: str
    joint_type: str
    rpy: List[float]
    xyz: List[float]
    axis: List[float]
    parent: int                      # index into UrdfModel.frames, -1 = base link
    has_collision: bool
    lower: Optional[float] = None
    upper: Optional[float] = None
    chain: List[int] = field(default_factory=list)   # frame indices base -> this frame (inclusive)


def _floats(text):
    return [float(tok) for tok in text.split()]


class UrdfModel:
    """Frame table of one robot."""

    def __init__(self, filepath):
        self.filepath = filepath
        robot = ElementTree.parse(filepath).getroot()
        links = {ln.attrib["name"]: ln for ln in robot.findall("link")}
        joints = robot.findall("joint")

        child_links = {j.find("child").attrib["link"] for j in joints}
        base = next(name for name in links if name not in child_links)
        self.base_link = base

        joints_of_parent = {}
        for j in joints:                                     # file order is preserved per parent
            joints_of_parent.setdefault(j.find("parent").attrib["link"], []).append(j)

        frames: List[Frame] = []
        queue = deque([(base, -1)])
        while queue:                                          # breadth first, like the reference
            link_name, parent_frame = queue.popleft()
            for j in joints_of_parent.get(link_name, []):
                jtype = j.attrib["type"]
                if jtype not in _TYPE_CODE:
                    raise NotImplementedError(f"joint type {jtype!r} of {j.attrib['name']!r} is not supported")
                origin = j.find("origin")
                rpy = _floats(origin.attrib.get("rpy", "0 0 0")) if origin is not None else [0.0, 0.0, 0.0]
                xyz = _floats(origin.attrib.get("xyz", "0 0 0")) if origin is not None else [0.0, 0.0, 0.0]
                axis_el = j.find("axis")
                if jtype == "fixed":
                    axis = [0.0, 0.0, 0.0]
                else:
                    axis = _floats(axis_el.attrib["xyz"]) if axis_el is not None else [1.0, 0.0, 0.0]
                child = j.find("child").attrib["link"]
                col = links[child].find("collision")
                # the reference tests the truth value of the <collision> element, i.e. "has children"
                has_collision = col is not None and len(col) > 0
                lim = j.find("limit")
                lower = float(lim.attrib["lower"]) if lim is not None and "lower" in lim.attrib else None
                upper = float(lim.attrib["upper"]) if lim is not None and "upper" in lim.attrib else None
                fr = Frame(j.attrib["name"], child, jtype, rpy, xyz, axis, parent_frame, has_collision, lower, upper)
                fr.chain = (frames[parent_frame].chain if parent_frame >= 0 else []) + [len(frames)]
                frames.append(fr)
                queue.append((child, len(frames) - 1))
        self.frames = frames

    # -- views -------------------------------------------------------------------------------
    @property
    def frame_names(self):
        return [f.name for f in self.frames]

    def index(self, frame_name):
        for i, f in enumerate(self.frames):
            if f.name == frame_name:
                return i
        raise KeyError(frame_name)

    def constant_transforms(self):
        """float32 [F,4,4]; rotation = R_x(roll) @ R_y(pitch) @ R_z(yaw) (reference order)."""
        out = np.zeros((len(self.frames), 4, 4), dtype=np.float32)
        for i, f in enumerate(self.frames):
            r, p, y = (np.float32(v) for v in f.rpy)
            cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
            one, zero = np.float32(1), np.float32(0)
            rx = np.array([[one, zero, zero], [zero, cr, -sr], [zero, sr, cr]], dtype=np.float32)
            ry = np.array([[cp, zero, sp], [zero, one, zero], [-sp, zero, cp]], dtype=np.float32)
            rz = np.array([[cy, -sy, zero], [sy, cy, zero], [zero, zero, one]], dtype=np.float32)
            out[i, :3, :3] = (rx @ ry) @ rz
            out[i, :3, 3] = np.asarray(f.xyz, dtype=np.float32)
            out[i, 3, 3] = 1.0
        return out

    def type_codes(self):
        return np.array([_TYPE_CODE[f.joint_type] for f in self.frames], dtype=np.int8)

    def axes(self):
        return np.array([f.axis for f in self.frames], dtype=np.float32).reshape(-1, 3)

    def parents(self):
        return np.array([f.parent for f in self.frames], dtype=np.int32)
