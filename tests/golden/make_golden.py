"""Generate the committed fixtures under tests/golden/.  Run in the BUILD container only:

    python tests/golden/make_golden.py

(a) urdf_frames.json -- produced by IMPORTING THE REFERENCE'S OWN ``helper/urdf_parsing.py`` (the one
    reference module that runs without TensorFlow/PyBullet) on the reference's own URDF files under
    /root/reference/urdf.  This pins frame order, joint types, rpy/xyz/axis and collision flags of
    both the oracle's and the product's URDF readers to the real reference.
(b) config*.npz / fk_*.npz -- seeded inputs with the outputs of the CPU oracle (oracle/rmp_oracle.py)
    in float32 (reference-faithful) and float64 (truth).  The reference itself (TensorFlow) cannot run
    here, so these are ORACLE outputs, not reference outputs.
(c) ref_*.npz come from a different script, run_reference_under_shim.py: the reference's own source files
    executed under oracle/tf_shim.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"

from oracle import harness as H                                   # noqa: E402
from oracle import rmp_oracle as O                                # noqa: E402
from riemannian_motion_policies_b200 import scenarios as S        # noqa: E402


def reference_frames():
    sys.path.insert(0, REFERENCE)
    from helper.urdf_parsing import UrdfTree                      # the reference's own parser
    out = {}
    for key, rel in (("panda", "urdf/franka_panda/panda.urdf"),
                     ("panda_wo_tool", "urdf/franka_panda/panda_wo_tool.urdf"),
                     ("two_joint", "urdf/TwoJointRobot_wo_fixedJoints.urdf"),
                     ("gantry", S.GANTRY_URDF)):          # this repo's synthetic test robot through the reference's parser
        tree = UrdfTree(os.path.join(REFERENCE, rel))
        frames = []
        for path in tree.get_backward_paths():
            e = tree.get_element_by_name(path[-1])
            frames.append(dict(name=e.name, link_name=e.link_name, joint_type=e.joint_type, rpy=e.rpy, xyz=e.xyz,
                               axis=e.axis, has_collision=bool(e.has_collision), path=path))
        out[key] = frames
    return out


def config_fixture(config, n, B):
    seed = S.SEEDS[config]
    if config == 1:
        q, qd, goal = S.sample_two_joint(B, seed)
        sph = None
    else:
        q, qd, goal = S.sample_panda_state(B, n, seed)
        O_ = S.N_SPHERES[config]
        if O_:
            fk = H.make_fkine(n, torch.float64)
            frames = S.collision_frames(fk)
            origins = np.stack([H.frame_origins(fk, torch.as_tensor(q[b]).double(), frames).numpy() for b in range(B)])
            sph = S.sample_spheres(B, O_, seed, origins)
        else:
            sph = None
    out32 = H.evaluate_vmap(config, n, q, qd, goal, sph, dtype=torch.float32)
    out64 = H.evaluate_vmap(config, n, q, qd, goal, sph, dtype=torch.float64)
    f64, M64 = H.combined_vmap(config, n, q, qd, goal, sph, dtype=torch.float64)
    d = dict(q=q, qd=qd, goal=goal, qdd32=out32, qdd64=out64, M64=M64, f64=f64)
    if sph is not None:
        d["spheres"] = sph
    return d


def fk_fixture(n, B, seed):
    fk = H.make_fkine(n, torch.float64)
    rng = np.random.RandomState(seed)
    if n == 2:
        q = rng.uniform(-np.pi, np.pi, size=(B, n))
    else:
        q = rng.uniform(S.PANDA_Q_LOW[:n], S.PANDA_Q_HIGH[:n], size=(B, n))
    qd = rng.uniform(-1, 1, size=(B, n))
    d = dict(q=q.astype(np.float32), qd=qd.astype(np.float32))
    for fi, frame in enumerate(fk.frame_names):
        xs, xds, Js, cs = [], [], [], []
        for b in range(B):
            x, xd, J, c = fk.differentiate(torch.as_tensor(d["q"][b]).double()[None], torch.as_tensor(d["qd"][b]).double()[None], frame)
            xs.append(x[0].numpy()); xds.append(xd[0].numpy()); Js.append(J[0].numpy()); cs.append(c[0].numpy())
        d[f"x_{fi}"], d[f"xd_{fi}"], d[f"J_{fi}"], d[f"c_{fi}"] = (np.stack(a) for a in (xs, xds, Js, cs))
    d["frame_names"] = np.array(fk.frame_names)
    return d


def main():
    if os.path.isdir(REFERENCE):
        with open(os.path.join(HERE, "urdf_frames.json"), "w") as fh:
            json.dump(reference_frames(), fh, indent=1)
        print("wrote urdf_frames.json from the reference parser")
    for config, n, B in ((1, 2, 64), (2, 7, 64), (2, 9, 32), (3, 7, 48), (3, 9, 16), (4, 7, 48), (5, 7, 32)):
        np.savez_compressed(os.path.join(HERE, f"config{config}_n{n}.npz"), **config_fixture(config, n, B))
        print("wrote config", config, "n", n)
    for n, B in ((2, 8), (7, 8), (9, 8)):
        np.savez_compressed(os.path.join(HERE, f"fk_n{n}.npz"), **fk_fixture(n, B, 100 + n))
        print("wrote fk", n)


if __name__ == "__main__":
    main()
