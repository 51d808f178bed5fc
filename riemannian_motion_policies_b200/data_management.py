"""Per-frame obstacle data store -- host-side mirror of the reference's ``data_management.py``.

Holds, for every frame, the closest-point pairs the simulation reports; task maps and leaves keep
*references* to these variables, so ``update()`` changes what the next control step sees
(reference: data_management.py:8-37, experiments/franka_panda/05_obstacle_avoidance.py:73-76).
"""
import numpy as np
import torch


class Variable:
    """Minimal stand-in for ``tf.Variable(shape=[None, ...])``: a re-assignable host tensor."""

    def __init__(self, initial_value):
        self._value = torch.as_tensor(initial_value, dtype=torch.float32)

    def assign(self, value):
        if isinstance(value, (list, tuple)):
            value = np.stack([np.asarray(v, dtype=np.float32) for v in value]) if len(value) else np.zeros((0,), np.float32)
        self._value = torch.as_tensor(value, dtype=torch.float32)
        return self

    def value(self):
        return self._value

    def numpy(self):
        return self._value.numpy()

    @property
    def shape(self):
        return self._value.shape

    def __len__(self):
        return self._value.shape[0]


class Datamanager:
    """reference: data_management.py:3-52."""

    KEYS = ('pos_on_link_in_base_frame', 'pos_on_obstacle_in_base_frame', 'normal_vec', 'distance',
            'relative_position')

    def __init__(self, fkine):
        self.fkine = fkine
        self.state = {
            frame_name: {
                'pos_on_link_in_base_frame': Variable(torch.zeros(0, 3)),
                'pos_on_obstacle_in_base_frame': Variable(torch.zeros(0, 3)),
                'normal_vec': Variable(torch.zeros(0, 3)),
                'distance': Variable(torch.zeros(0)),
                'relative_position': Variable(torch.zeros(0, 3)),
            } for frame_name in self.fkine.frame_names
        }

    def __getitem__(self, key):
        return self.state[key]

    def update(self, q, distance_data):
        """distance_data: iterable of (frame_name, pos_on_link[3], pos_on_obstacle[3], normal[3],
        distance, description) tuples (reference: simulation.py:462-484).  Frames without an entry
        keep their previous values, like the reference (data_management.py:23-27)."""
        by_frame = {}
        for d in distance_data:
            by_frame.setdefault(d[0], []).append(d)
        for frame_name in self.fkine.frame_names:
            rows = by_frame.get(frame_name)
            if not rows:
                continue
            st = self.state[frame_name]
            link = np.stack([np.asarray(d[1], dtype=np.float32) for d in rows])
            st['pos_on_link_in_base_frame'].assign(link)
            st['pos_on_obstacle_in_base_frame'].assign(np.stack([np.asarray(d[2], dtype=np.float32) for d in rows]))
            st['normal_vec'].assign(np.stack([np.asarray(d[3], dtype=np.float32) for d in rows]))
            st['distance'].assign(np.asarray([d[4] for d in rows], dtype=np.float32))
            st['relative_position'].assign(self._get_relative_pos(link, frame_name, q))

    def _get_relative_pos(self, pos_on_link_in_base_frame, frame_name, q):
        """Link points expressed in the joint frame (reference: data_management.py:44-52); one FK
        call per frame instead of one per pair."""
        T = self.fkine.forward(np.asarray(q, dtype=np.float32)[None], frame_name)[0]
        T = T.cpu() if T.is_cuda else T
        R, p = T[:3, :3], T[:3, 3]
        rel = torch.as_tensor(pos_on_link_in_base_frame) - p
        return rel @ R                      # (R^T rel^T)^T
