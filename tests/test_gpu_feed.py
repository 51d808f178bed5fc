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


def _segment_distance_bruteforce(P1, Q1, P2, Q2, n=401):
    """min |x - y| over two segments by dense sampling + one refinement (independent of the kernel's closed form)."""
    s = np.linspace(0, 1, n)
    X = P1[None] + s[:, None] * (Q1 - P1)[None]
    Y = P2[None] + s[:, None] * (Q2 - P2)[None]
    D = np.linalg.norm(X[:, None, :] - Y[None, :, :], axis=-1)
    i, j = np.unravel_index(np.argmin(D), D.shape)
    lo = lambda k: max(s[max(k - 1, 0)], 0.0)
    hi = lambda k: min(s[min(k + 1, n - 1)], 1.0)
    s1, s2 = np.linspace(lo(i), hi(i), n), np.linspace(lo(j), hi(j), n)
    X = P1[None] + s1[:, None] * (Q1 - P1)[None]
    Y = P2[None] + s2[:, None] * (Q2 - P2)[None]
    return np.linalg.norm(X[:, None, :] - Y[None, :, :], axis=-1).min()


def test_link_capsules_against_bruteforce(ns):
    """Control geometry = a capsule riding on each frame (the reference asks PyBullet for the closest points between
    the LINK's collision shape and the obstacle, simulation.py:462-484): reported points lie on the two surfaces,
    the normal joins them, and the distance equals a brute-force segment-segment minimum minus the radii."""
    n = 9
    fk = product_fkine(ns, n)
    rng = np.random.RandomState(5)
    frames = S.collision_frames(fk)
    links = {fr: (rng.uniform(-0.05, 0.05, 3), rng.uniform(-0.12, 0.12, 3), float(rng.uniform(0.02, 0.06))) for fr in frames[:-2]}
    links[frames[-2]] = (np.array([0.0, 0.0, 0.03]), np.array([0.0, 0.0, 0.03]), 0.04)      # a sphere; the last frame: origin
    feed = ObstacleFeed(fk, link_capsules=links)
    B, O, C = 4, 3, 4
    q, _, _ = S.sample_panda_state(B, n, seed=6)
    spheres = np.concatenate([rng.uniform(-0.8, 0.8, size=(B, O, 3)), rng.uniform(0.02, 0.1, size=(B, O, 1))], -1).astype(np.float32)
    a = rng.uniform(-0.8, 0.8, size=(B, C, 3))
    b = a + rng.uniform(-0.3, 0.3, size=(B, C, 3))
    b[:, 0] = a[:, 0] + np.array([0.0, 0.0, 0.4])                                            # one vertical cylinder
    capsules = np.concatenate([a, b, rng.uniform(0.02, 0.05, size=(B, C, 1)), np.zeros((B, C, 1))], -1).astype(np.float32)
    pairs, aux = feed.closest_points(q, spheres, capsules)
    pairs, aux = pairs.cpu().numpy().astype(np.float64), aux.cpu().numpy().astype(np.float64)
    K = O + C
    ofk = H.make_fkine(n, torch.float64)
    for e in range(B):
        for fi, frame in enumerate(feed.frames):
            T = ofk.forward(torch.as_tensor(q[e]).double()[None], frame)[0].numpy()
            la, lb, lr = links.get(frame, (np.zeros(3), np.zeros(3), 0.0))
            P1, Q1 = T[:3, 3] + T[:3, :3] @ la, T[:3, 3] + T[:3, :3] @ lb
            for o in range(K):
                row, ar = pairs[e, fi * K + o], aux[e, fi * K + o]
                if o < O:
                    P2 = Q2 = spheres[e, o, :3].astype(np.float64)
                    rad = float(spheres[e, o, 3])
                else:
                    cc = capsules[e, o - O].astype(np.float64)
                    P2, Q2, rad = cc[0:3], cc[3:6], cc[6]
                want = _segment_distance_bruteforce(P1, Q1, P2, Q2) - lr - rad
                assert abs(ar[0] - want) < 2e-5, (frame, o, ar[0], want)
                on_link, on_obst, nrm = row[0:3], row[3:6], ar[1:4]
                assert abs(np.linalg.norm(nrm) - 1) < 1e-5
                np.testing.assert_allclose(on_link - on_obst, ar[0] * nrm, atol=5e-6)         # the normal joins the points
                # the points lie on the two surfaces: distance to the axis segments equals the radii
                assert abs(_segment_distance_bruteforce(on_link, on_link, P1, Q1) - lr) < 2e-5
                assert abs(_segment_distance_bruteforce(on_obst, on_obst, P2, Q2) - rad) < 2e-5


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
