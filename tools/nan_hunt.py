"""Diagnostic: run the closed loop of tests/test_gpu_properties.py::test_closed_loop_reaches_the_goal_without_penetration
one control step at a time and dump the state (q, qd, goal, spheres, frame origins / velocities) of the first
environment whose command turns non-finite -> gpurun_out/nan_hunt.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_common import closed_loop_scene                            # noqa: E402
from riemannian_motion_policies_b200 import scenarios as S          # noqa: E402

N = 7
ns = S.product_namespace()
dev = torch.device("cuda")
fk = ns.UrdfForwardKinematic(S.PANDA_WO_TOOL_URDF, S.PANDA_ORDER_7)
Bc, O_, dt, every = 512, 8, 0.01, 10
q0, qd0, goal, sph = closed_loop_scene(Bc, O_, seed=7)
q, qd = torch.as_tensor(q0, device=dev), torch.as_tensor(qd0, device=dev)
spheres = torch.as_tensor(sph, device=dev)
goals = torch.as_tensor(goal, device=dev).reshape(Bc, 1, 3).contiguous()
core = S.build_config3(ns, fk, [0.5, 0.0, 0.5], N, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
tree = core.compile(N, goal_leaves=["attractor"])
qdd = torch.empty(Bc, N, device=dev)
frames = S.collision_frames(fk)
for step in range(600):
    qp, qdp = q.clone(), qd.clone()
    tree.rollout(q, qd, qdd, dt, every, every, goals=goals, spheres=spheres)
    bad = ~torch.isfinite(qdd).all(dim=1) | ~torch.isfinite(q).all(dim=1)
    if bool(bad.any()):
        e = int(torch.nonzero(bad)[0])
        print("first non-finite at control step", step, "env", e, "qdd", qdd[e].tolist())
        print("q", qp[e].tolist(), "qd", qdp[e].tolist())
        origins = torch.stack([fk.forward(qp[e:e + 1], fr)[:, :3, 3] for fr in frames], dim=1)[0]
        d = torch.linalg.norm(origins[:, None, :] - spheres[e][None, :, :3], dim=-1) - spheres[e][None, :, 3]
        print("surface distances min", d.min().item(), "max |qd|", qdp[e].abs().max().item())
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez(os.path.join(ROOT, "gpurun_out", "nan_hunt.npz"), q=qp[e].cpu().numpy(), qd=qdp[e].cpu().numpy(),
                 goal=goals[e].cpu().numpy(), spheres=spheres[e].cpu().numpy(), origins=origins.cpu().numpy(),
                 d=d.cpu().numpy(), step=step, env=e)
        # same state through a plain step, all pairs and early-out
        for eo in (False, True):
            tree.set_early_out(eo)
            out = torch.empty(1, N, device=dev)
            tree.step(qp[e:e + 1].contiguous(), qdp[e:e + 1].contiguous(), out, goals=goals[e:e + 1].contiguous(),
                      spheres=spheres[e:e + 1].contiguous())
            print("plain step early_out", eo, out[0].tolist())
        break
else:
    print("no non-finite command in 600 control steps; max |qd|", qd.abs().max().item())
