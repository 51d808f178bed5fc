"""On-GPU obstacle feed: the step *before* the control step.

Replaces ``Simulation.calculate_distances`` (reference: simulation.py:462-484, PyBullet closest-point
queries between link geometry and obstacles) with primitive geometry on both sides: obstacles are spheres and
capsules (the experiments' cylinders, experiments/franka_panda/06_cluttered_environment.py:39-52); every listed
collision frame carries one capsule fixed in the frame (``link_capsules``) -- or, by default, just its origin.  The output uses the reference's own wire format, so the unchanged
``Datamanager.update(q, distance_data)`` path (data_management.py:22-37) consumes it; for batches the
pair rows can be handed to ``CompiledTree.step(..., pairs=...)`` directly.
"""
import numpy as np
import torch

from . import _native
from ._tensor import current_stream_ptr, require_cuda, to_device


class ObstacleFeed:
    def __init__(self, fkine, frames=None, link_capsules=None):
        """link_capsules: {frame_name: (a[3], b[3], radius)} in FRAME coordinates -- the link geometry riding on the
        frame (a == b: a sphere); frames not in the mapping use their origin as control point."""
        self.fkine = fkine
        if frames is None:
            frames = [name for name, has in zip(fkine.frame_names, fkine.has_collision) if has]
        self.frames = list(frames)
        self._idx = np.array([fkine.frame_index(f) for f in self.frames], dtype=np.int32)
        self._links = None
        if link_capsules:
            unknown = set(link_capsules) - set(self.frames)
            if unknown:
                raise KeyError(f"link_capsules for frames that are not listed: {sorted(unknown)}")
            self._links = np.zeros((len(self.frames), 8), dtype=np.float32)
            for i, frame in enumerate(self.frames):
                if frame in link_capsules:
                    a, b, radius = link_capsules[frame]
                    self._links[i, 0:3], self._links[i, 3:6], self._links[i, 6] = a, b, radius

    def closest_points(self, q, spheres=None, capsules=None):
        """q [B,n]; spheres [B,O,4]; capsules [B,C,8] -> pairs [B, F*K, 8], aux [B, F*K, 4] (CUDA tensors);
        row order: frame-major (the order of ``self.frames``), spheres before capsules."""
        dev = require_cuda()
        qt = to_device(q, dev).reshape(-1, self.fkine.n_joints)
        B = qt.shape[0]
        sp = None if spheres is None else to_device(spheres, dev).reshape(B, -1, 4)
        cp = None if capsules is None else to_device(capsules, dev).reshape(B, -1, 8)
        O = 0 if sp is None else sp.shape[1]
        C = 0 if cp is None else cp.shape[1]
        K = O + C
        pairs = torch.empty(B, len(self.frames) * K, _native.PAIR_FLOATS, device=dev)
        aux = torch.empty(B, len(self.frames) * K, 4, device=dev)
        if B and K:
            _native.check(_native.lib().rmp2_obstacle_feed(
                self.fkine._handle, self._idx.ctypes.data, None if self._links is None else self._links.ctypes.data,
                len(self.frames), B, qt.data_ptr(),
                None if sp is None else sp.data_ptr(), O, None if cp is None else cp.data_ptr(), C,
                pairs.data_ptr(), aux.data_ptr(), current_stream_ptr(dev)))
        return pairs, aux

    def state(self, q, spheres=None, capsules=None):
        """Single environment, the reference's ``distance_data``: a list of tuples
        (frame_name, pos_on_link[3], pos_on_obstacle[3], normal_vec[3], distance, description)."""
        q = np.asarray(q, dtype=np.float32).reshape(1, -1)
        sp = None if spheres is None else np.asarray(spheres, dtype=np.float32).reshape(1, -1, 4)
        cp = None if capsules is None else np.asarray(capsules, dtype=np.float32).reshape(1, -1, 8)
        pairs, aux = self.closest_points(q, sp, cp)
        pairs, aux = pairs[0].cpu().numpy(), aux[0].cpu().numpy()
        K = pairs.shape[0] // max(1, len(self.frames))
        out = []
        for fi, frame in enumerate(self.frames):
            for o in range(K):
                row, a = pairs[fi * K + o], aux[fi * K + o]
                out.append((frame, row[0:3].copy(), row[3:6].copy(), a[1:4].copy(), float(a[0]),
                            f'{frame} to obstacle {o}'))
        return out
