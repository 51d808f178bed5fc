"""Parity of the CUDA control step (through the C ABI) with the CPU oracle -- the tests proper."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from gpu_common import (REL_TOL, assert_parity, config_sens, make_inputs, product_core, product_evaluate, product_fkine,
                        rel_err)
from oracle import harness as H
from riemannian_motion_policies_b200 import scenarios as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ns(native_lib):
    return S.product_namespace()


@pytest.mark.parametrize("config,n", [(1, 2), (2, 7), (2, 9), (3, 7), (3, 9), (4, 7), (5, 7)])
def test_golden_fixtures(ns, config, n):
    """Committed vectors (tests/golden, oracle f32 + f64 outputs)."""
    g = np.load(os.path.join(GOLDEN, f"config{config}_n{n}.npz"))
    sph = g["spheres"] if "spheres" in g else None
    got = product_evaluate(ns, config, n, g["q"], g["qd"], g["goal"], sph)
    stats = assert_parity(got, g["qdd32"], g["qdd64"], g["M64"], n, label=f"golden config{config} n{n}",
                          max_excluded=0.10 if config == 4 else 0.05, sens=config_sens(config, n, g["q"], g["qd"], g["goal"], sph))
    print(f"config{config} n{n}: {stats}")


@pytest.mark.parametrize("config,n", [(1, 2), (2, 7), (2, 9), (3, 7), (3, 9), (4, 7), (5, 7), (6, 9)])
def test_reference_source_vectors(ns, config, n):
    """tests/golden/ref_*.npz: outputs of the REFERENCE'S OWN source files run under the TensorFlow-API shim
    (tests/golden/run_reference_under_shim.py).  The kernel must match them like it matches the oracle."""
    g = np.load(os.path.join(GOLDEN, f"ref_config{config}_n{n}.npz"))
    sph = g["spheres"] if "spheres" in g else None
    got = product_evaluate(ns, config, n, g["q"], g["qd"], g["goal"], sph)
    ref64 = H.evaluate_vmap(config, n, g["q"], g["qd"], g["goal"], sph, dtype=torch.float64)
    _, M64 = H.combined_vmap(config, n, g["q"], g["qd"], g["goal"], sph, dtype=torch.float64)
    stats = assert_parity(got, g["qdd_ref"], ref64, M64, n, label=f"reference-source config{config} n{n}",
                          max_excluded=0.10 if config == 4 else 0.05, sens=config_sens(config, n, g["q"], g["qd"], g["goal"], sph))
    print(f"reference-source config{config} n{n}: {stats}")


def test_reference_source_vectors_v1(ns):
    """The v1 CollisionAvoidance tree of the reference (two-joint experiment 05) on the same distance_data."""
    g = np.load(os.path.join(GOLDEN, "ref_v1_two_joint.npz"))
    fk = product_fkine(ns, 2)
    frames = list(g["frames"])
    got = []
    for b in range(g["q"].shape[0]):
        q, qd, goal, rows = g["q"][b], g["qd"][b], g["goal"][b], g["distance_rows"][b]
        distance_data = [(frames[i], rows[i, 0:3], rows[i, 3:6], rows[i, 6:9], rows[i, 9], "golden") for i in range(len(frames))]
        dm = ns.Datamanager(fk)
        core = ns.RmpCore()
        core.add_rmp(ns.TargetPolicy(alpha=0.1, beta=0.1, c=0.1, goal=goal, name="target",
                                     taskmap=S.ee_position_taskmap(ns, fk, "link_23")))
        for frame in fk.frame_names:
            tm = ns.chain_taskmaps([ns.TaskmapByForwardKinematic(fk, frame),
                                    ns.TaskmapRelative4x4(relative_pos=dm[frame]["relative_position"]),
                                    ns.TaskmapFrom4x4ToPosition()])
            core.add_rmp(ns.CollisionAvoidance(d=dm[frame]["distance"], vec=dm[frame]["normal_vec"],
                                               eta_rep=0.1 * np.e, nu_rep=0.3, eta_damp=1, nu_damp=0.3, r=1.1, c=1e5,
                                               taskmap=tm, name=f"collision_avoidance_for_{frame}"))
        dm.update(q, distance_data)
        got.append(core.evaluate(q, qd).numpy())
    from test_reference_golden import v1_oracle
    ref64 = np.stack([v1_oracle(g, b, torch.float64) for b in range(g["q"].shape[0])])
    assert_parity(np.stack(got), g["qdd_ref"], ref64, label="reference-source v1 CollisionAvoidance")


@pytest.mark.parametrize("config,n,B", [(1, 2, 1000), (2, 7, 4096), (3, 7, 2048), (4, 7, 1024), (4, 9, 512), (5, 7, 1024), (5, 9, 256),
                                        (6, 9, 512)])
def test_seeded_batches_against_oracle(ns, config, n, B):
    """Same seeded inputs through the oracle (vmap, f32 and f64) and the kernel."""
    q, qd, goal, sph = make_inputs(config, n, B)
    ref32 = H.evaluate_vmap(config, n, q, qd, goal, sph, dtype=torch.float32)
    ref64 = H.evaluate_vmap(config, n, q, qd, goal, sph, dtype=torch.float64)
    _, M64 = H.combined_vmap(config, n, q, qd, goal, sph, dtype=torch.float64)
    got = product_evaluate(ns, config, n, q, qd, goal, sph)
    # config 4 with the two finger joints (n = 9): the fingers only carry the joint-limit metric, which puts a singular
    # value within 4x of the pinv cutoff in 37 % of these environments (measured on a B200: 191 of 512 excluded, the
    # other 321 all pass) -- the truncation is discontinuous there for every float32 implementation
    max_excluded = 0.45 if (config, n) == (4, 9) else 0.10 if config == 4 else 0.05
    assert np.isfinite(got).all()
    stats = assert_parity(got, ref32, ref64, M64, n, label=f"config{config} n{n} B{B}",
                          max_excluded=max_excluded, sens=config_sens(config, n, q, qd, goal, sph))
    print(f"config{config} n{n} B{B}: {stats}")
    if config in (2, 3, 5):      # mostly well-conditioned trees: the strict 1e-5 bar holds almost everywhere
        assert stats["frac_strict"] > 0.95, stats
    # config 6 = the synthetic gantry arm: revolute joints about general axes (the non-z Rodrigues branch of
    # chain_advance), prismatic joints along x / y / z, multi-axis rpy constants, three kinematic branchings (chain
    # state slots, > 48 KB of dynamic shared memory), scrambled joint order, and an orientation leaf on the Euler
    # task map (RMP2_SPACE_FRAME_EULER)


@pytest.mark.parametrize("config", [4, 5])
def test_large_seeded_batches_against_committed_oracle_outputs(ns, config):
    """4096 seeded environments of the two 64-sphere trees against oracle outputs computed in the build container
    (tests/golden/make_parity_fixtures.py -> parity_config*_n7.npz: the oracle needs minutes for these).  The inputs
    are regenerated from the seed and checked against the fixture's digest."""
    import hashlib
    g = np.load(os.path.join(GOLDEN, f"parity_config{config}_n7.npz"))
    B = int(g["B"])
    q, qd, goal, sph = make_inputs(config, 7, B)
    h = hashlib.sha256()
    for a in (q, qd, goal, sph):
        h.update(np.ascontiguousarray(a).tobytes())
    assert h.hexdigest() == str(g["digest"]), "seeded inputs differ from the ones the fixture was computed for"
    got = product_evaluate(ns, config, 7, q, qd, goal, sph)
    stats = assert_parity(got, g["ref32"], g["ref64"], n=7, s64=g["s64"], label=f"config{config} B{B} (fixture)",
                          max_excluded=0.10 if config == 4 else 0.05, sens=g["sens"])
    if config == 5:
        assert stats["frac_strict"] >= 0.97, stats


def test_reference_style_single_env_call(ns):
    """`core.evaluate(q, qd).numpy()` with 1-D numpy inputs, as the experiments call it
    (reference: experiments/two_joint_robot/01_target_rmp_only.py:55)."""
    fk = product_fkine(ns, 2)
    q, qd, goal = S.sample_two_joint(20, seed=7)
    outs = []
    for b in range(20):
        core = S.build_config1(ns, fk, goal[b])
        out = core.evaluate(q[b], qd[b])
        assert isinstance(out, torch.Tensor) and not out.is_cuda and out.shape == (2,)
        outs.append(out.numpy())
    ref32 = H.evaluate_loop(1, 2, q, qd, goal)
    ref64 = H.evaluate_loop(1, 2, q, qd, goal, dtype=torch.float64)
    _, M64 = H.combined_vmap(1, 2, q, qd, goal, dtype=torch.float64)
    assert_parity(np.stack(outs), ref32, ref64, M64, 2, label="single-env calls", max_excluded=0.2,
                  sens=config_sens(1, 2, q, qd, goal, None))
    # reassigning the goal attribute is picked up at the next step (06_cluttered_environment.py:142)
    core = S.build_config1(ns, fk, goal[0])
    a = core.evaluate(q[0], qd[0]).numpy()
    core.rmps['target'].goal = np.array(goal[1])
    b_ = core.evaluate(q[0], qd[0]).numpy()
    ref = H.evaluate_loop(1, 2, q[0:1], qd[0:1], goal[1:2])[0]
    assert rel_err(b_, ref) <= REL_TOL and not np.allclose(a, b_)


def test_datamanager_pairs_feed_matches_oracle(ns):
    """The reference's own obstacle wire format: per-frame closest-point pairs through Datamanager
    (data_management.py:22-37), ragged K per frame including K = 0, n = 9 panda."""
    n = 9
    fk = product_fkine(ns, n)
    ofk = H.make_fkine(n)
    rng = np.random.RandomState(11)
    q, qd, goal = S.sample_panda_state(6, n, seed=12)
    frames = S.collision_frames(fk)
    for b in range(6):
        origins = H.frame_origins(ofk, torch.as_tensor(q[b]), frames).numpy()
        distance_data = []
        for k, frame in enumerate(frames):
            K = [0, 1, 3, 7, 2, 5, 4, 1, 6, 2][k] if b % 2 == 0 else rng.randint(0, 5)
            for _ in range(K):
                on_link = origins[k] + rng.uniform(-0.05, 0.05, size=3)
                direction = rng.normal(size=3)
                direction /= np.linalg.norm(direction)
                dist = rng.uniform(0.04, 0.45)
                on_obst = on_link + dist * direction
                distance_data.append((frame, on_link.astype(np.float32), on_obst.astype(np.float32),
                                      (-direction).astype(np.float32), np.float32(dist), 'synthetic'))
        # product: Datamanager variables captured by reference inside the task maps
        dm = ns.Datamanager(fk)
        core = S.build_config3(ns, fk, goal[b], n, lambda fr: ns.TaskmapJointFrame4x4ToDistance(
            dm[fr]['pos_on_link_in_base_frame'], dm[fr]['pos_on_obstacle_in_base_frame']))
        dm.update(q[b], distance_data)
        got = core.evaluate(q[b], qd[b]).numpy()
        # oracle: same tuples, same tree builder
        out = {}
        for dtype in (torch.float32, torch.float64):
            ons = H.namespace(dtype)
            fko = H.make_fkine(n, dtype)
            pts = {fr: ([d[1] for d in distance_data if d[0] == fr], [d[2] for d in distance_data if d[0] == fr]) for fr in frames}
            tm_for = lambda fr: ons.TaskmapJointFrame4x4ToDistance(
                torch.tensor(np.array(pts[fr][0]).reshape(-1, 3)), torch.tensor(np.array(pts[fr][1]).reshape(-1, 3)))
            ocore = S.build_config3(ons, fko, torch.as_tensor(goal[b]), n, tm_for)
            out[dtype] = ocore.evaluate(torch.as_tensor(q[b]), torch.as_tensor(qd[b])).numpy()
        e32, e64 = rel_err(got, out[torch.float32]), rel_err(got, out[torch.float64])
        yard = rel_err(out[torch.float32], out[torch.float64])
        assert e32 <= REL_TOL or e64 <= max(REL_TOL, 2 * yard), (b, e32, e64, yard)


def test_v1_collision_avoidance_two_joint(ns):
    """The v1 obstacle path of experiments/two_joint_robot/05_obstacle_avoidance.py:44-61:
    chain [FK(frame), TaskmapRelative4x4(relative_position), TaskmapFrom4x4ToPosition] + CollisionAvoidance
    on every frame, fed through Datamanager (distance, normal_vec, relative_position), plus TargetPolicy."""
    from oracle import rmp_oracle as O
    fk = product_fkine(ns, 2)
    rng = np.random.RandomState(21)
    q_all, qd_all, goal_all = S.sample_two_joint(12, seed=22)
    got, ref32, ref64 = [], [], []
    for b in range(12):
        q, qd, goal = q_all[b], qd_all[b], goal_all[b]
        distance_data = []
        for frame in fk.frame_names:
            for _ in range(rng.randint(0, 4)):                      # ragged, possibly empty
                T = fk.forward(q[None], frame)[0].numpy()
                on_link = T[:3, 3] + T[:3, :3] @ rng.uniform(-0.3, 0.3, size=3)
                direction = rng.normal(size=3)
                direction /= np.linalg.norm(direction)
                dist = rng.uniform(0.05, 1.3)                      # some beyond r = 1.1
                distance_data.append((frame, on_link.astype(np.float32), (on_link - dist * direction).astype(np.float32),
                                      direction.astype(np.float32), np.float32(dist), 'synthetic'))

        def build(m, fkine, dm):
            core = m.RmpCore()
            core.add_rmp(m.TargetPolicy(alpha=0.1, beta=0.1, c=0.1, goal=goal, name='target',
                                        taskmap=S.ee_position_taskmap(m, fkine, 'link_23')))
            for frame in fkine.frame_names:
                tm = m.chain_taskmaps([m.TaskmapByForwardKinematic(fkine, frame),
                                       m.TaskmapRelative4x4(relative_pos=dm[frame]['relative_position']),
                                       m.TaskmapFrom4x4ToPosition()])
                core.add_rmp(m.CollisionAvoidance(d=dm[frame]['distance'], vec=dm[frame]['normal_vec'],
                                                  eta_rep=0.1 * np.e, nu_rep=0.3, eta_damp=1, nu_damp=0.3, r=1.1, c=1e5,
                                                  taskmap=tm, name=f'collision_avoidance_for_{frame}'))
            return core

        dm = ns.Datamanager(fk)
        core = build(ns, fk, dm)
        dm.update(q, distance_data)
        got.append(core.evaluate(q, qd).numpy())
        # oracle side: same tuples; relative_position computed the reference's way (data_management.py:44-52)
        for dtype, sink in ((torch.float32, ref32), (torch.float64, ref64)):
            ons = H.namespace(dtype)
            fko = H.make_fkine(2, dtype)
            odm = {}
            for frame in fko.frame_names:
                rows = [d for d in distance_data if d[0] == frame]
                T = fko.forward(torch.as_tensor(q)[None], frame)[0]
                rel = [T[:3, :3].T @ (torch.as_tensor(d[1]).to(dtype) - T[:3, 3]) for d in rows]
                odm[frame] = {'relative_position': torch.stack(rel) if rel else torch.zeros(0, 3, dtype=dtype),
                              'distance': torch.tensor([float(d[4]) for d in rows], dtype=dtype),
                              'normal_vec': torch.tensor(np.array([d[3] for d in rows]).reshape(-1, 3), dtype=dtype)}
            sink.append(build(ons, fko, odm).evaluate(torch.as_tensor(q), torch.as_tensor(qd)).numpy())
    assert_parity(np.stack(got), np.stack(ref32), np.stack(ref64), label="v1 CollisionAvoidance two-joint")


def test_edge_cases(ns):
    n = 7
    fk = product_fkine(ns, n)
    core = product_core(ns, 3, n, fk)
    dev = torch.device("cuda")
    # empty batch
    out = core.evaluate(torch.zeros(0, n, device=dev), torch.zeros(0, n, device=dev),
                        goals=torch.zeros(0, 3, device=dev), spheres=torch.zeros(0, 16, 4, device=dev))
    assert out.shape == (0, n)
    # ragged sizes: B not a multiple of the warp / block, O not a multiple of 8 (non-TMA path), O = 0
    for B, O_ in ((1, 16), (33, 16), (129, 5), (200, 0), (77, 40), (50, 64), (40, 100)):   # O > 64: the chunked early-out
        q, qd, goal = S.sample_panda_state(B, n, seed=20 + B)
        ofk = H.make_fkine(n, torch.float64)
        frames = S.collision_frames(ofk)
        origins = torch.func.vmap(lambda qq: H.frame_origins(ofk, qq, frames))(torch.as_tensor(q).double()).numpy()
        sph = S.sample_spheres(B, O_, 30 + B, origins) if O_ else None
        got = product_evaluate(ns, 3, n, q, qd, goal, sph, fkine=fk, core=core)
        if sph is None:        # oracle with zero obstacle leaves active: park one sphere far away
            sph_o = np.tile(np.array([[[5.0, 5.0, 5.0, 0.05]]], np.float32), (B, 1, 1))
        else:
            sph_o = sph
        ref32 = H.evaluate_vmap(3, n, q, qd, goal, sph_o, dtype=torch.float32)
        ref64 = H.evaluate_vmap(3, n, q, qd, goal, sph_o, dtype=torch.float64)
        assert_parity(got, ref32, ref64, label=f"edge B{B} O{O_}", sens=config_sens(3, n, q, qd, goal, sph_o))


def test_tma_and_direct_paths_agree(ns):
    """The TMA-staged sphere path and the plain global-load path run the same arithmetic."""
    from riemannian_motion_policies_b200 import _native
    n, B = 7, 1000
    q, qd, goal, sph = make_inputs(4, n, B)
    fk = product_fkine(ns, n)
    core = product_core(ns, 4, n, fk)
    dev = torch.device("cuda")
    tree = core.compile(n, goal_leaves=["attractor"])
    args = [torch.as_tensor(a, device=dev) for a in (q, qd)]
    goals = torch.as_tensor(goal, device=dev).reshape(B, 1, 3).contiguous()
    spheres = torch.as_tensor(sph, device=dev)
    out = {}
    for tma in (1, 0):
        tree.set_option(_native.OPT_TMA, tma)
        qdd = torch.empty(B, n, device=dev)
        tree.step(args[0], args[1], qdd, goals=goals, spheres=spheres)
        out[tma] = qdd.cpu().numpy()
    np.testing.assert_array_equal(out[1], out[0])
    # a sphere pointer that is not 16-byte aligned is refused (rows are read as float4), not mis-read
    flat = torch.zeros(B * 64 * 4 + 1, device=dev)
    shifted = flat[1:].view(B, 64, 4)
    with pytest.raises(ValueError, match="16-byte aligned"):
        tree.step(args[0], args[1], torch.empty(B, n, device=dev), goals=goals, spheres=shifted)


@pytest.mark.parametrize("config", [2, 4, 5])
def test_split_and_fused_resolve_agree(ns, config):
    """RMP2_OPT_SPLIT_RESOLVE: the direct resolve inside the step kernel (default) or as its own kernel behind the
    (M, f) scratch -- same arithmetic, same fallback list, bit-identical results."""
    from riemannian_motion_policies_b200 import _native
    n, B = 7, 3000
    q, qd, goal, sph = make_inputs(config, n, B)
    fk = product_fkine(ns, n)
    core = product_core(ns, config, n, fk)
    dev = torch.device("cuda")
    tree = core.compile(n, goal_leaves=["target" if config == 2 else "attractor"])
    tq, tqd = torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev)
    goals = torch.as_tensor(goal, device=dev).reshape(B, 1, 3).contiguous()
    spheres = None if sph is None else torch.as_tensor(sph, device=dev)
    out = {}
    for split in (0, 1):
        tree.set_option(_native.OPT_SPLIT_RESOLVE, split)
        qdd = torch.empty(B, n, device=dev)
        tree.step(tq, tqd, qdd, goals=goals, spheres=spheres)
        out[split] = qdd.cpu().numpy()
    np.testing.assert_array_equal(out[0], out[1])


def test_early_out_is_exact(ns):
    """Skipping pairs beyond the metric radius (library default) changes nothing: they contribute
    exactly zero in the reference (rmp2.py:194)."""
    n, B = 7, 4096
    q, qd, goal, sph = make_inputs(5, n, B)
    fk = product_fkine(ns, n)
    core = product_core(ns, 5, n, fk)
    dev = torch.device("cuda")
    args = [torch.as_tensor(a, device=dev) for a in (q, qd)]
    goals = torch.as_tensor(goal, device=dev).reshape(B, 1, 3).contiguous()
    spheres = torch.as_tensor(sph, device=dev)
    tree = core.compile(n, goal_leaves=["attractor"])
    out = {}
    for flag in (True, False):
        tree.set_early_out(flag)
        qdd = torch.empty(B, n, device=dev)
        tree.step(args[0], args[1], qdd, goals=goals, spheres=spheres)
        out[flag] = qdd.cpu().numpy()
    tree.set_early_out(True)
    np.testing.assert_array_equal(out[True], out[False])


@pytest.mark.parametrize("config,n", [(4, 7), (5, 7), (5, 9)])
def test_merged_coincident_leaves_equal_the_unmerged_tree(ns, config, n):
    """RMP2_OPT_MERGE_COINCIDENT (library default): one pair loop for the obstacle leaves that share their control point
    (Panda: joint6 on joint5), none for control points that cannot move (joint1, joint2: J = 0).  M1 + M1 = 2 M1 is exact and the pulled-back terms are the same numbers,
    only the order of the sum over leaves changes: the two commands agree to float32 rounding amplified by the
    conditioning of the resolve, both variants of the pair kernel, generic and specialised kernels alike."""
    B = 2048
    q, qd, goal, sph = make_inputs(config, n, B)
    fk = product_fkine(ns, n)
    core = product_core(ns, config, n, fk)
    dev = torch.device("cuda")
    qt, qdt = (torch.as_tensor(a, device=dev) for a in (q, qd))
    goals = torch.as_tensor(goal, device=dev).reshape(B, 1, 3).contiguous()
    spheres = torch.as_tensor(sph, device=dev)
    tree = core.compile(n, goal_leaves=["attractor"])
    leaves, slots = tree.obstacle_slots()
    assert slots == leaves - 3          # joint6 rides on joint5; joint1 / joint2 cannot move (J = 0): no pair loop at all

    def run():
        qdd = torch.empty(B, n, device=dev)
        tree.step(qt, qdt, qdd, goals=goals, spheres=spheres)
        return qdd.cpu().numpy().astype(np.float64)

    out = {}
    for merged in (True, False):
        tree.set_merge_coincident(merged)
        for early in (True, False):
            tree.set_early_out(early)
            out[merged, early] = run()
        np.testing.assert_array_equal(out[merged, True], out[merged, False])     # the early-out stays exact
    tree.specialize()
    spec_off = run()
    tree.set_merge_coincident(True)
    assert tree.obstacle_slots() == (leaves, slots) and tree.specialized_seconds() is not None
    spec_on = run()
    tree.set_early_out(True)
    np.testing.assert_array_equal(spec_off, out[False, False])
    np.testing.assert_array_equal(spec_on, out[True, False])
    ref64 = H.evaluate_vmap(config, n, q, qd, goal, sph, dtype=torch.float64)
    scale = np.linalg.norm(ref64, axis=1)
    e_merge = np.linalg.norm(out[True, False] - out[False, False], axis=1) / scale
    e_on = np.linalg.norm(out[True, False] - ref64, axis=1) / scale
    e_off = np.linalg.norm(out[False, False] - ref64, axis=1) / scale
    print(f"config {config} n={n}: merged vs unmerged median {np.median(e_merge):.2e} q99 {np.quantile(e_merge, .99):.2e}; "
          f"vs f64 oracle: merged median {np.median(e_on):.2e}, unmerged {np.median(e_off):.2e}")
    # the difference between the two is float32 rounding of the same sums: no larger than either one's distance from
    # the float64 oracle (medians), and the merged tree is as close to the oracle as the unmerged one
    assert np.median(e_merge) <= 2 * max(np.median(e_off), 1e-7)
    assert np.median(e_on) <= 1.25 * np.median(e_off) + 1e-7
    # (config 4 is rank deficient: environments at the pinv cutoff flip with any change of rounding -- its tail is
    # compared at the 90 % quantile, config 5's at 99 %)
    tail = 0.90 if config == 4 else 0.99
    assert np.quantile(e_on, tail) <= 2 * np.quantile(e_off, tail) + 1e-6


@pytest.mark.parametrize("n_leaves", [1, 3])
def test_early_out_with_few_obstacle_leaves(ns, n_leaves):
    """One or three obstacle leaves instead of eight: the pair kernel's blocks shrink to 32 / 96 threads (32 environments
    per tile); the re-dealing early-out must still be bit-identical to the all-pairs variant and match the oracle."""
    n, B, O_ = 7, 1500, 64
    q, qd, goal = S.sample_panda_state(B, n, seed=61)
    fk = product_fkine(ns, n)
    ofk = H.make_fkine(n, torch.float64)
    frames = S.collision_frames(ofk)[-n_leaves:]
    origins = torch.func.vmap(lambda qq: H.frame_origins(ofk, qq, frames))(torch.as_tensor(q).double()).numpy()
    sph = S.sample_spheres(B, O_, 62, origins)

    def build(ns_, fk_, g, tm_for):
        core = ns_.RmpCore()
        core.add_rmp(S.target_attractor(ns_, fk_, g))
        core.add_rmp(ns_.JointDamping(accel_d_gain=1, metric_scalar=0.005, inertia=0.3))
        for fr in frames:
            core.add_rmp(S.obstacle_leaf(ns_, ns_.chain_taskmaps([ns_.TaskmapByForwardKinematic(fk_, fr), tm_for(fr)]), fr))
        return core

    dev = torch.device("cuda")
    tree = build(ns, fk, goal[0], lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance()).compile(n, goal_leaves=["attractor"])
    tq, tqd = torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev)
    goals = torch.as_tensor(goal, device=dev).reshape(B, 1, 3).contiguous()
    spheres = torch.as_tensor(sph, device=dev)
    out = {}
    for flag in (True, False):
        tree.set_early_out(flag)
        qdd = torch.empty(B, n, device=dev)
        tree.step(tq, tqd, qdd, goals=goals, spheres=spheres)
        out[flag] = qdd.cpu().numpy()
    np.testing.assert_array_equal(out[True], out[False])

    def oracle(dtype, inputs=None):
        ons, fko = H.namespace(dtype), H.make_fkine(n, dtype)
        q_, qd_, goal_, sph_ = inputs if inputs is not None else (q, qd, goal, sph)

        def one(q1, qd1, g1, s):
            q1, qd1, g1, s = q1.to(dtype), qd1.to(dtype), g1.to(dtype), s.to(dtype)
            org = H.frame_origins(fko, q1, frames)
            r = org[:, None, :] - s[None, :, :3]
            on_obst = s[None, :, :3] + s[None, :, 3:4] * r / torch.linalg.norm(r, dim=-1, keepdim=True)
            on_link = org[:, None, :].expand_as(on_obst)
            idx = {fr: i for i, fr in enumerate(frames)}
            return build(ons, fko, g1, lambda fr: ons.TaskmapJointFrame4x4ToDistance(on_link[idx[fr]], on_obst[idx[fr]])).evaluate(q1, qd1)

        return torch.func.vmap(one)(torch.as_tensor(q_), torch.as_tensor(qd_), torch.as_tensor(goal_), torch.as_tensor(sph_)).numpy()

    sens = lambda idx: H.sensitivity(lambda *arrs: oracle(torch.float64, inputs=arrs), [q[idx], qd[idx], goal[idx], sph[idx]])
    assert_parity(out[True], oracle(torch.float32), oracle(torch.float64), label=f"{n_leaves} obstacle leaves", sens=sens)


def test_joint_subset_pads_the_kernel_width(ns):
    """n = 5 controllable joints of the 7-joint arm (kernel instantiated for 7): joints outside `order`
    read q = 0 like the reference (kinematics.py:197,218-219); padded rows/cols stay zero."""
    from oracle import rmp_oracle as O
    order = S.PANDA_ORDER_7[:5]
    fk = ns.UrdfForwardKinematic(S.PANDA_WO_TOOL_URDF, order)
    rng = np.random.RandomState(3)
    B = 64
    q = rng.uniform(S.PANDA_Q_LOW[:5], S.PANDA_Q_HIGH[:5], size=(B, 5)).astype(np.float32)
    qd = rng.uniform(-0.3, 0.3, size=(B, 5)).astype(np.float32)
    goal = rng.uniform([0.3, -0.7, 0.3], [0.7, 0.7, 0.7], size=(B, 3)).astype(np.float32)
    core = S.build_config2(ns, fk, goal[0], 5)
    core.add_rmp(ns.JointDamping(accel_d_gain=1, metric_scalar=0.005, inertia=0.3))
    got = core.evaluate(torch.as_tensor(q).cuda(), torch.as_tensor(qd).cuda(), goals=torch.as_tensor(goal).cuda()).cpu().numpy()
    for dtype, tol in ((torch.float32, REL_TOL),):
        ons = H.namespace(dtype)
        fko = O.UrdfForwardKinematic(S.PANDA_WO_TOOL_URDF, order, dtype=dtype)
        ref = []
        for b in range(B):
            oc = S.build_config2(ons, fko, torch.as_tensor(goal[b]), 5)
            oc.add_rmp(ons.JointDamping(accel_d_gain=1, metric_scalar=0.005, inertia=0.3))
            ref.append(oc.evaluate(torch.as_tensor(q[b]), torch.as_tensor(qd[b])).numpy())
        assert rel_err(got, np.stack(ref)).max() <= tol


def _tree_with_obstacle_gains(ns_, fk, goal, n, tm_for, **gains):
    """target attractor + joint damping + one ObstacleAvoidance leaf per collision frame with the given gains."""
    base = dict(margin=0., damping_gain=50, damping_std_dev=0.04, damping_robustness_eps=0.01,
                damping_velocity_gate_length_scale=0.01, repulsion_gain=800, repulsion_std_dev=0.01,
                metric_modulation_radius=0.5, metric_scalar=1, metric_exploder_std_dev=0.02, metric_exploder_eps=0.001)
    base.update(gains)
    core = ns_.RmpCore()
    core.add_rmp(S.target_attractor(ns_, fk, goal))
    core.add_rmp(ns_.JointDamping(accel_d_gain=1, metric_scalar=0.005, inertia=0.3))
    for frame in S.collision_frames(fk):
        tm = ns_.chain_taskmaps([ns_.TaskmapByForwardKinematic(fk, frame), tm_for(frame)])
        core.add_rmp(ns_.ObstacleAvoidance(taskmap=tm, name=f"oa_{frame}", **base))
    return core


@pytest.mark.parametrize("gains", [
    dict(margin=0.05),
    dict(margin=0.02, metric_modulation_radius=0.3, metric_scalar=2.5),
    dict(metric_scalar=-0.5, repulsion_gain=-100.0),          # signs the reference accepts too
    dict(damping_velocity_gate_length_scale=0.002),           # exp(xdot / l_v) overflows for receding pairs
])
def test_obstacle_leaf_parameter_variants(ns, gains):
    """The packed pair loop folds margin, radius and metric_scalar into its coefficients on the host
    (fill_sphere_row): non-default gains must still match the oracle, which evaluates rmp2.py:184-196 as written."""
    n, B, O_ = 7, 512, 16
    q, qd, goal = S.sample_panda_state(B, n, seed=77)
    ofk = H.make_fkine(n, torch.float64)
    frames = S.collision_frames(ofk)
    origins = torch.func.vmap(lambda qq: H.frame_origins(ofk, qq, frames))(torch.as_tensor(q).double()).numpy()
    sph = S.sample_spheres(B, O_, 78, origins)
    fk = product_fkine(ns, n)
    core = _tree_with_obstacle_gains(ns, fk, goal[0], n, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance(), **gains)
    dev = torch.device("cuda")
    got = core.evaluate(torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev),
                        goals=torch.as_tensor(goal, device=dev), spheres=torch.as_tensor(sph, device=dev)).cpu().numpy()

    def oracle(dtype, combine=False, inputs=None):
        ons = H.namespace(dtype)
        fko = H.make_fkine(n, dtype)
        q_, qd_, goal_, sph_ = inputs if inputs is not None else (q, qd, goal, sph)

        def one(q1, qd1, goal1, s):
            q1, qd1, goal1, s = q1.to(dtype), qd1.to(dtype), goal1.to(dtype), s.to(dtype)
            org = H.frame_origins(fko, q1, frames)
            r = org[:, None, :] - s[None, :, :3]
            on_obst = s[None, :, :3] + s[None, :, 3:4] * r / torch.linalg.norm(r, dim=-1, keepdim=True)
            on_link = org[:, None, :].expand_as(on_obst)
            idx = {fr: i for i, fr in enumerate(frames)}
            oc = _tree_with_obstacle_gains(ons, fko, goal1, n,
                                           lambda fr: ons.TaskmapJointFrame4x4ToDistance(on_link[idx[fr]], on_obst[idx[fr]]),
                                           **gains)
            return oc.combine(q1, qd1) if combine else oc.evaluate(q1, qd1)

        res = torch.func.vmap(one)(torch.as_tensor(q_), torch.as_tensor(qd_), torch.as_tensor(goal_), torch.as_tensor(sph_))
        return tuple(r.numpy() for r in res) if combine else res.numpy()

    f64, M64 = oracle(torch.float64, combine=True)
    sens = lambda idx: np.maximum(
        H.sensitivity(lambda *arrs: oracle(torch.float64, inputs=arrs), [q[idx], qd[idx], goal[idx], sph[idx]]),
        H.metric_conditioning(M64[idx]))
    stats = assert_parity(got, oracle(torch.float32), oracle(torch.float64), M64, n, label=f"obstacle gains {gains}", sens=sens)
    print(gains, stats)


def test_inert_and_degenerate_spheres(ns):
    """metric_scalar = 0 and spheres beyond every metric radius contribute exactly nothing (bitwise);
    a frame origin inside a sphere or exactly on its centre stays finite."""
    n, B, O_ = 7, 256, 12
    q, qd, goal = S.sample_panda_state(B, n, seed=91)
    fk = product_fkine(ns, n)
    dev = torch.device("cuda")
    tq, tqd, tg = (torch.as_tensor(a, device=dev) for a in (q, qd, goal))
    tm_for = lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance()
    ofk = H.make_fkine(n, torch.float64)
    frames = S.collision_frames(ofk)
    origins = torch.func.vmap(lambda qq: H.frame_origins(ofk, qq, frames))(torch.as_tensor(q).double()).numpy()
    near = torch.as_tensor(S.sample_spheres(B, O_, 92, origins), device=dev)
    far = near.clone()
    far[..., :3] += 50.0
    plain = ns.RmpCore()
    plain.add_rmp(S.target_attractor(ns, fk, goal[0]))
    plain.add_rmp(ns.JointDamping(accel_d_gain=1, metric_scalar=0.005, inertia=0.3))
    want = plain.evaluate(tq, tqd, goals=tg).cpu().numpy()
    inert = _tree_with_obstacle_gains(ns, fk, goal[0], n, tm_for, metric_scalar=0.0)
    np.testing.assert_array_equal(inert.evaluate(tq, tqd, goals=tg, spheres=near).cpu().numpy(), want)
    live = _tree_with_obstacle_gains(ns, fk, goal[0], n, tm_for)
    np.testing.assert_array_equal(live.evaluate(tq, tqd, goals=tg, spheres=far).cpu().numpy(), want)
    # sphere 0 swallows the last collision frame's origin, sphere 1 is centred exactly on the first one's
    bad = near.clone()
    bad[:, 0, :3] = torch.as_tensor(origins[:, -1], device=dev, dtype=torch.float32) + 0.01
    bad[:, 0, 3] = 0.05
    bad[:, 1, :3] = torch.as_tensor(origins[:, 0], device=dev, dtype=torch.float32)
    out = live.evaluate(tq, tqd, goals=tg, spheres=bad)
    assert torch.isfinite(out).all()


def test_direct_solve_and_jacobi_lanes_mix(ns):
    """Trees with an isotropic metric leaf solve by plain QR where the matrix is provably clear of the pinv
    cutoff and fall back to the Jacobi sweeps otherwise, lane by lane.  A weak isotropic weight puts
    sigma_min/sigma_max ~ 1e-4 (no truncation, but around the rigorous test's threshold), so both kinds of
    lanes share warps: every environment must match the oracle, and must not depend on its neighbours."""
    from oracle import rmp_oracle as O
    n, B = 7, 512
    q, qd, goal = S.sample_panda_state(B, n, seed=55)
    fk = product_fkine(ns, n)

    def build(ns_, fk_, g):
        core = ns_.RmpCore()
        core.add_rmp(ns_.TargetPolicy(alpha=0.1, beta=1, c=0.1, goal=g, name="target", taskmap=S.ee_position_taskmap(ns_, fk_)))
        core.add_rmp(ns_.ConfigurationSpaceBiasing(gamma_p=0.01, gamma_d=0.1, q0=S.NULLSPACE_Q0_9[:n], name="bias", w=1e-4))
        return core

    dev = torch.device("cuda")
    core = build(ns, fk, goal[0])
    tq, tqd, tg = (torch.as_tensor(a, device=dev) for a in (q, qd, goal))
    got = core.evaluate(tq, tqd, goals=tg).cpu().numpy()

    def oracle(dtype, combine=False, inputs=None):
        ons = H.namespace(dtype)
        fko = H.make_fkine(n, dtype)
        q_, qd_, goal_ = inputs if inputs is not None else (q, qd, goal)

        def one(q1, qd1, g1):
            oc = build(ons, fko, g1.to(dtype))
            return oc.combine(q1.to(dtype), qd1.to(dtype)) if combine else oc.evaluate(q1.to(dtype), qd1.to(dtype))

        res = torch.func.vmap(one)(torch.as_tensor(q_), torch.as_tensor(qd_), torch.as_tensor(goal_))
        return tuple(r.numpy() for r in res) if combine else res.numpy()

    f64, M64 = oracle(torch.float64, combine=True)
    sens = lambda idx: np.maximum(H.sensitivity(lambda *arrs: oracle(torch.float64, inputs=arrs), [q[idx], qd[idx], goal[idx]]),
                                  H.metric_conditioning(M64[idx]))
    s = np.linalg.svd(M64, compute_uv=False)
    ratio = s[:, -1] / s[:, 0]
    assert (ratio > 4 * 10 * n * np.finfo(np.float32).eps).all()          # nothing is truncated in this batch
    assert (ratio < 2.4e-4).any() and (ratio > 2.4e-4).any()              # ... but it straddles the direct-solve test
    stats = assert_parity(got, oracle(torch.float32), oracle(torch.float64), M64, n, label="direct/jacobi mix", sens=sens)
    print(stats, "sigma ratio range", ratio.min(), ratio.max())
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).to(dev)
    again = core.evaluate(tq[perm], tqd[perm], goals=tg[perm]).cpu().numpy()
    np.testing.assert_array_equal(again, got[perm.cpu().numpy()])


@pytest.mark.parametrize("config,n", [(1, 2), (2, 7), (3, 7), (3, 9), (4, 7), (5, 7)])
def test_specialized_kernels_match(ns, config, n):
    """rmp2_tree_specialize: the frames / step kernels rebuilt by NVRTC for one tree (tables as compile-time
    constants, loops unrolled) compute what the generic, table-interpreting kernels compute -- same source,
    so nearly always the same bits -- and meet the oracle parity criterion on their own."""
    B = 2048 if config != 1 else 512
    q, qd, goal, sph = make_inputs(config, n, B)
    fk = product_fkine(ns, n)
    core = product_core(ns, config, n, fk)
    dev = torch.device("cuda")
    tq, tqd, tg = (torch.as_tensor(a, device=dev) for a in (q, qd, goal))
    ts = None if sph is None else torch.as_tensor(sph, device=dev)
    goal_leaves = ["target"] if config in (1, 2) else ["attractor"]
    tree = core.compile(n, goal_leaves=goal_leaves)
    goals = tg.reshape(B, 1, 3).contiguous()
    generic = torch.empty(B, n, device=dev)
    tree.step(tq, tqd, generic, goals=goals, spheres=ts)
    assert tree.specialized_seconds() is None
    seconds = tree.specialize()
    assert seconds is not None and seconds > 0
    special = torch.empty(B, n, device=dev)
    tree.step(tq, tqd, special, goals=goals, spheres=ts)
    same = (generic == special).all(dim=1).float().mean().item()
    err = rel_err(special.cpu().numpy(), generic.cpu().numpy())
    print(f"config{config} n{n}: NVRTC {seconds:.1f} s, bit-identical environments {same:.3f}, median rel diff {np.median(err):.2e}")
    ref32 = H.evaluate_vmap(config, n, q, qd, goal, sph, dtype=torch.float32)
    ref64 = H.evaluate_vmap(config, n, q, qd, goal, sph, dtype=torch.float64)
    _, M64 = H.combined_vmap(config, n, q, qd, goal, sph, dtype=torch.float64)
    assert_parity(special.cpu().numpy(), ref32, ref64, M64, n, label=f"specialized config{config} n{n}",
                  max_excluded=0.10 if config == 4 else 0.05, sens=config_sens(config, n, q, qd, goal, sph))
    # the fused-resolve kernel (small batches) is specialised too
    small = torch.empty(64, n, device=dev)
    tree.step(tq[:64].contiguous(), tqd[:64].contiguous(), small, goals=goals[:64].contiguous(),
              spheres=None if ts is None else ts[:64].contiguous())
    np.testing.assert_allclose(small.cpu().numpy(), special[:64].cpu().numpy(), rtol=2e-3, atol=1e-5)


def test_leaf_update_keeps_the_specialization(ns):
    """The tables are compile-time constants of the specialised kernels, so changing a leaf parameter (reference
    idiom: ``target_rmp.goal = ...``, 06_cluttered_environment.py:142) rebuilds them: the next step gives the new
    goal's answer and still runs the specialised kernels."""
    n, B = 7, 256
    q, qd, goal, _ = make_inputs(2, n, B)
    fk = product_fkine(ns, n)
    core = product_core(ns, 2, n, fk)
    dev = torch.device("cuda")
    tq, tqd = torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev)
    tree = core.compile(n)
    tree.specialize()
    a = core.evaluate(tq, tqd).cpu().numpy()
    assert core.compile(n).specialized_seconds() is not None
    core.rmps["target"].goal = [0.3, 0.2, 0.6]
    b = core.evaluate(tq, tqd).cpu().numpy()
    assert core.compile(n).specialized_seconds() is not None
    assert np.abs(a - b).max() > 1e-4
    fresh = product_core(ns, 2, n, fk)                      # generic kernels, same goal
    fresh.rmps["target"].goal = [0.3, 0.2, 0.6]
    np.testing.assert_allclose(fresh.evaluate(tq, tqd).cpu().numpy(), b, rtol=1e-5, atol=1e-6)
