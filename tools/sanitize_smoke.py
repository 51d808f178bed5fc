"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_common import make_inputs, product_evaluate, product_fkine       # noqa: E402
from riemannian_motion_policies_b200 import scenarios as S                # noqa: E402
from riemannian_motion_policies_b200.obstacle_feed import ObstacleFeed    # noqa: E402

ns = S.product_namespace()
for config, n, B in ((1, 2, 77), (2, 9, 100), (3, 7, 333), (4, 7, 200), (5, 9, 130), (5, 7, 40000)):
    q, qd, goal, sph = make_inputs(config, n, B)
    out = product_evaluate(ns, config, n, q, qd, goal, sph)
    assert np.isfinite(out).all()
    # odd sphere counts -> non-TMA path
    if sph is not None:
        out = product_evaluate(ns, config, n, q, qd, goal, sph[:, :13].copy())
        assert np.isfinite(out).all()
fk = product_fkine(ns, 9)
q, qd, goal = S.sample_panda_state(50, 9, 1)
fk.differentiate(torch.as_tensor(q).cuda(), torch.as_tensor(qd).cuda(), "panda_finger_joint2")
feed = ObstacleFeed(fk)
feed.closest_points(q, spheres=np.random.rand(50, 3, 4).astype(np.float32), capsules=np.random.rand(50, 2, 8).astype(np.float32))
leaf = ns.ObstacleAvoidance(0., 50, 0.04, 0.01, 0.01, 800, 0.01, 0.5, 1, 0.02, 0.001, ns.IdentityTaskmap(), 'o')
leaf.evaluate(np.random.rand(64, 1).astype(np.float32), np.random.rand(64, 1).astype(np.float32))
# host path + rollout
fk7 = product_fkine(ns, 7)
core = S.build_config5(ns, fk7, [0.5, 0, 0.5], 7, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
tree = core.compile(7, goal_leaves=["attractor"])
q, qd, goal, sph = make_inputs(5, 7, 3000)
qh, qdh, gh, sh = (torch.as_tensor(a) for a in (q, qd, goal.reshape(-1, 1, 3), sph))
qdd = torch.empty(3000, 7)
tree.step_host(qh, qdh, qdd, goals=gh.contiguous(), spheres=sh)
qc, qdc = qh.cuda(), qdh.cuda()
tree.rollout(qc, qdc, torch.empty(3000, 7, device="cuda"), 0.01, 20, 10, goals=gh.cuda().contiguous(), spheres=sh.cuda())
torch.cuda.synchronize()
print("sanitize smoke ok")
