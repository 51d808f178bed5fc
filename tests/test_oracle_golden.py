"""The oracle against the committed golden vectors (tests/golden/*.npz, made by make_golden.py):
guards the checker itself against drift (torch version, refactors).  float32 outputs must reproduce
to rounding, float64 'truth' outputs tightly."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import harness as H


def _rel(a, b):
    return np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1e-30)


@pytest.mark.parametrize("config,n", [(1, 2), (2, 7), (3, 7), (4, 7), (5, 7)])
def test_oracle_reproduces_golden(config, n):
    g = np.load(os.path.join(GOLDEN, f"config{config}_n{n}.npz"))
    B = min(16, g["q"].shape[0])
    sph = g["spheres"][:B] if "spheres" in g else None
    out64 = H.evaluate_vmap(config, n, g["q"][:B], g["qd"][:B], g["goal"][:B], sph, dtype=torch.float64)
    assert _rel(out64, g["qdd64"][:B]).max() < 1e-9
    out32 = H.evaluate_vmap(config, n, g["q"][:B], g["qd"][:B], g["goal"][:B], sph, dtype=torch.float32)
    # float32 reproduces up to reduction-order effects, scaled by how ill-conditioned the env is
    err_ref = _rel(g["qdd32"][:B], g["qdd64"][:B])
    assert (_rel(out32, g["qdd64"][:B]) <= np.maximum(1e-5, 4 * err_ref)).all()


def test_loop_and_vmap_modes_agree():
    g = np.load(os.path.join(GOLDEN, "config3_n7.npz"))
    a = H.evaluate_loop(3, 7, g["q"][:2], g["qd"][:2], g["goal"][:2], g["spheres"][:2])
    assert _rel(a, g["qdd64"][:2]).max() < 1e-5


def test_pinv_cutoff_is_float32_in_both_modes():
    """rcond = 10 * n * eps32 even when arithmetic is float64."""
    from oracle import rmp_oracle as O
    M = torch.diag(torch.tensor([1.0, 1e-3, 5e-6, 1e-9], dtype=torch.float64))
    P = O.tf_pinv(M)
    assert P[2, 2] == pytest.approx(2e5) and P[3, 3] == 0.0     # cutoff = 10*4*eps32 = 4.8e-6
    P32 = O.tf_pinv(M.float())
    assert P32[2, 2].item() == pytest.approx(2e5, rel=1e-5) and P32[3, 3].item() == 0.0


@pytest.mark.parametrize("n,groups", [(7, [("panda_joint1", "panda_joint2"), ("panda_joint5", "panda_joint6")]),
                                      (9, [("panda_joint1", "panda_joint2"), ("panda_joint5", "panda_joint6")])])
def test_obstacle_leaves_on_coincident_frame_origins_pull_back_identically(n, groups):
    """What RMP2_OPT_MERGE_COINCIDENT rests on, checked on the restated reference (float64, autodiff derivatives): the
    distance map differentiates through the frame origin only (taskmap.py:124-128), so ObstacleAvoidance leaves with
    equal gains on frames whose origins coincide for every q (a zero constant translation behind a revolute joint:
    Panda joint2 on joint1, joint6 on joint5) contribute the same pulled-back (f, M) -- the reference computes it twice
    and adds; the CUDA tree compiler runs one pair loop and doubles its sums.  Frames that do not coincide differ."""
    from riemannian_motion_policies_b200 import scenarios as S
    dtype = torch.float64
    ns, fk = H.namespace(dtype), H.make_fkine(n, dtype)
    frames = S.collision_frames(fk)
    q, qd, goal = S.sample_panda_state(3, n, seed=11)
    rng = np.random.RandomState(12)
    for e in range(3):
        sph = torch.as_tensor(np.concatenate([rng.uniform([-0.6, -0.6, 0.1], [0.6, 0.6, 1.0], size=(6, 3)),
                                              rng.uniform(0.03, 0.08, size=(6, 1))], -1), dtype=dtype)
        qe, qde = torch.as_tensor(q[e], dtype=dtype), torch.as_tensor(qd[e], dtype=dtype)
        origins = H.frame_origins(fk, qe, frames)
        r = origins[:, None, :] - sph[None, :, :3]
        on_obst = sph[None, :, :3] + sph[None, :, 3:4] * r / torch.linalg.norm(r, dim=-1, keepdim=True)
        on_link = origins[:, None, :].expand_as(on_obst)
        idx = {fr: i for i, fr in enumerate(frames)}
        core = S.build_config4(ns, fk, torch.as_tensor(goal[e], dtype=dtype), n,
                               lambda fr: ns.TaskmapJointFrame4x4ToDistance(on_link[idx[fr]], on_obst[idx[fr]]))
        pulled = {}
        for fr in frames:
            f, M = core._calculate_rmp(core.rmps[f"collision_avoidance_for_{fr}"], qe, qde)
            pulled[fr] = (f.sum(0).numpy(), M.sum(0).numpy())
        scale = max(np.abs(M).max() for _, M in pulled.values()) + 1e-30
        for a, b in groups:
            np.testing.assert_allclose(pulled[a][0], pulled[b][0], rtol=0, atol=1e-12 * max(1.0, np.abs(pulled[a][0]).max()))
            np.testing.assert_allclose(pulled[a][1], pulled[b][1], rtol=0, atol=1e-12 * scale)
        # ... and it is the coincidence that does it: joint3 sits 0.316 m away from joint2
        assert np.abs(pulled["panda_joint3"][1] - pulled["panda_joint2"][1]).max() > 1e-6 * scale \
            or np.abs(pulled["panda_joint3"][1]).max() == 0.0
