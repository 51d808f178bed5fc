"""On-GPU obstacle feed (the step before the hot path): closest points on spheres / capsules in the
reference's distance_data wire format, then through the UNCHANGED Datamanager -> task map -> leaf path."""
import numpy as np
import pytest
import torch

from gpu_common import product_fkine, rel_err
from oracle import harness as H
from riemannian_motion_policies_b200 import scenarios as S
from riemannian_motion_policies_b200.obstacle_feed import ObstacleFeed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ns(native_lib):
    return S.product_namespace()


def _closest_on_capsule(p, a, b, r):
    u = b - a
    t = np.clip(np.dot(p - a, u) / np.dot(u, u), 0.0, 1.0)
    c = a + t * u
    n = (p - c) / np.linalg.norm(p - c)
    return c + r * n, n, np.linalg.norm(p - c) - r


def test_closest_points_against_numpy(ns):
    n = 9
    fk = product_fkine(ns, n)
    feed = ObstacleFeed(fk)
    assert feed.frames == S.collision_frames(fk)
    rng = np.random.RandomState(0)
    B, O, C = 5, 3, 4
    q, _, _ = S.sample_panda_state(B, n, seed=1)
    spheres = np.concatenate([rng.uniform(-0.8, 0.8, size=(B, O, 3)), rng.uniform(0.02, 0.1, size=(B, O, 1))], -1).astype(np.float32)
    a = rng.uniform(-0.8, 0.8, size=(B, C, 3))
    b = a + rng.uniform(-0.3, 0.3, size=(B, C, 3))
    capsules = np.concatenate([a, b, rng.uniform(0.02, 0.05, size=(B, C, 1)), np.zeros((B, C, 1))], -1).astype(np.float32)
    pairs, aux = feed.closest_points(q, spheres, capsules)
    pairs, aux = pairs.cpu().numpy(), aux.cpu().numpy()
    K = O + C
    assert pairs.shape == (B, len(feed.frames) * K, 8) and aux.shape == (B, len(feed.frames) * K, 4)
    ofk = H.make_fkine(n, torch.float64)
    for e in range(B):
        origins = H.frame_origins(ofk, torch.as_tensor(q[e]).double(), feed.frames).numpy()
        for fi in range(len(feed.frames)):
            for o in range(K):
                row, ar = pairs[e, fi * K + o], aux[e, fi * K + o]
                if o < O:
                    c, r = spheres[e, o, :3].astype(np.float64), float(spheres[e, o, 3])
                    nrm = (origins[fi] - c) / np.linalg.norm(origins[fi] - c)
                    on_obst, dist = c + r * nrm, np.linalg.norm(origins[fi] - c) - r
                else:
                    cc = capsules[e, o - O].astype(np.float64)
                    on_obst, nrm, dist = _closest_on_capsule(origins[fi], cc[0:3], cc[3:6], cc[6])
                np.testing.assert_allclose(row[0:3], origins[fi], atol=2e-6)
                np.testing.assert_allclose(row[3:6], on_obst, atol=3e-6)
                np.testing.assert_allclose(ar[1:4], nrm, atol=2e-5)
                assert abs(ar[0] - dist) < 3e-6


def test_feed_through_datamanager_equals_sphere_path(ns):
    """distance_data from the feed -> Datamanager.update -> [FK, JointFrame4x4ToDistance] leaves gives the
    same command as the fused sphere path (spheres=...) of the very same tree."""
    n = 9
    fk = product_fkine(ns, n)
    feed = ObstacleFeed(fk)
    q, qd, goal = S.sample_panda_state(4, n, seed=3)
    rng = np.random.RandomState(4)
    for e in range(4):
        spheres = np.concatenate([rng.uniform([-0.6, -0.6, 0.2], [0.6, 0.6, 1.0], size=(6, 3)),
                                  rng.uniform(0.02, 0.06, size=(6, 1))], -1).astype(np.float32)
        distance_data = feed.state(q[e], spheres=spheres)
        assert len(distance_data) == len(feed.frames) * 6 and distance_data[0][0] == feed.frames[0]
        dm = ns.Datamanager(fk)
        core_pairs = S.build_config3(ns, fk, goal[e], n, lambda fr: ns.TaskmapJointFrame4x4ToDistance(
            dm[fr]['pos_on_link_in_base_frame'], dm[fr]['pos_on_obstacle_in_base_frame']))
        dm.update(q[e], distance_data)
        a = core_pairs.evaluate(q[e], qd[e]).numpy()
        core_sph = S.build_config3(ns, fk, goal[e], n, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
        b = core_sph.evaluate(q[e], qd[e], spheres=spheres[None]).numpy()
        assert rel_err(a[None], b[None])[0] < 2e-5
