"""Full-size checks (BASELINE.json sizes: 1,048,576 environments, 64 spheres) through properties that
do not need the oracle at that size, plus an oracle spot check on a random subset."""
import numpy as np
import pytest
import torch

from gpu_common import assert_parity, config_sens
from oracle import harness as H
from riemannian_motion_policies_b200 import scenarios as S

pytestmark = pytest.mark.gpu

B_FULL = 1 << 20
N = 7


@pytest.fixture(scope="module")
def world(native_lib):
    ns = S.product_namespace()
    dev = torch.device("cuda")
    fk = ns.UrdfForwardKinematic(S.PANDA_WO_TOOL_URDF, S.PANDA_ORDER_7)
    core = S.build_config4(ns, fk, [0.5, 0.0, 0.5], N, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
    tree = core.compile(N, goal_leaves=["attractor"])
    q, qd, goal, spheres = S.synth_inputs_device(fk, N, B_FULL, 64, 1, seed=3, device=dev)
    goals = goal.reshape(B_FULL, 1, 3).contiguous()
    qdd = torch.empty(B_FULL, N, device=dev)
    tree.step(q, qd, qdd, goals=goals, spheres=spheres[0])
    torch.cuda.synchronize()
    return dict(ns=ns, fk=fk, core=core, tree=tree, q=q, qd=qd, goals=goals, spheres=spheres[0], qdd=qdd, dev=dev)


def test_full_size_output_is_finite_and_deterministic(world):
    w = world
    assert torch.isfinite(w["qdd"]).all()
    again = torch.empty_like(w["qdd"])
    w["tree"].step(w["q"], w["qd"], again, goals=w["goals"], spheres=w["spheres"])
    assert torch.equal(again, w["qdd"])


def test_environments_are_independent(world):
    """Permutation equivariance and chunk invariance: an environment's result does not depend on its
    position in the batch, nor on what else is in the batch (blocks, tails, TMA tiles, chunking)."""
    w = world
    g = torch.Generator(device=w["dev"])
    g.manual_seed(0)
    perm = torch.randperm(B_FULL, generator=g, device=w["dev"])
    out = torch.empty_like(w["qdd"])
    w["tree"].step(w["q"][perm].contiguous(), w["qd"][perm].contiguous(), out, goals=w["goals"][perm].contiguous(),
                   spheres=w["spheres"][perm].contiguous())
    assert torch.equal(out, w["qdd"][perm])
    for lo, hi in ((0, 1), (5, 5 + 33), (1000, 1000 + 65537), (B_FULL - 12345, B_FULL)):
        part = torch.empty(hi - lo, N, device=w["dev"])
        w["tree"].step(w["q"][lo:hi].contiguous(), w["qd"][lo:hi].contiguous(), part,
                       goals=w["goals"][lo:hi].contiguous(), spheres=w["spheres"][lo:hi].contiguous())
        assert torch.equal(part, w["qdd"][lo:hi])


def test_batches_beyond_one_internal_chunk(world):
    """rmp2_step processes at most 2^20 environments per internal chunk (bounded scratch); a batch of
    2^20 + 12345 environments must give the same rows as the pieces it is made of."""
    w = world
    extra = 12345
    cat = lambda t: torch.cat([t, t[:extra]], dim=0).contiguous()
    out = torch.empty(B_FULL + extra, N, device=w["dev"])
    w["tree"].step(cat(w["q"]), cat(w["qd"]), out, goals=cat(w["goals"]), spheres=cat(w["spheres"]))
    assert torch.equal(out[:B_FULL], w["qdd"])
    assert torch.equal(out[B_FULL:], w["qdd"][:extra])


def test_full_size_subset_against_oracle(world):
    w = world
    rng = np.random.RandomState(1)
    idx = torch.as_tensor(np.sort(rng.choice(B_FULL, size=512, replace=False)), device=w["dev"])
    q, qd = w["q"][idx].cpu().numpy(), w["qd"][idx].cpu().numpy()
    goal, sph = w["goals"][idx, 0].cpu().numpy(), w["spheres"][idx].cpu().numpy()
    ref32 = H.evaluate_vmap(4, N, q, qd, goal, sph, dtype=torch.float32)
    ref64 = H.evaluate_vmap(4, N, q, qd, goal, sph, dtype=torch.float64)
    _, M64 = H.combined_vmap(4, N, q, qd, goal, sph, dtype=torch.float64)
    stats = assert_parity(w["qdd"][idx].cpu().numpy(), ref32, ref64, M64, N, label="1M-env subset", max_excluded=0.10,
                          sens=config_sens(4, N, q, qd, goal, sph))
    print(stats)


def test_host_buffer_path_matches_device_path(world):
    """rmp2_step_host (pinned host tensors, chunked 3-stream pipeline) == rmp2_step."""
    w = world
    Bh = 200_000                         # three 64k chunks + a ragged one
    pin = lambda t: t[:Bh].cpu().pin_memory()
    qdd_h = torch.empty(Bh, N).pin_memory()
    w["tree"].step_host(pin(w["q"]), pin(w["qd"]), qdd_h, goals=pin(w["goals"]), spheres=pin(w["spheres"]))
    assert torch.equal(qdd_h, w["qdd"][:Bh].cpu())
    # pageable (non-pinned) host memory also works
    qdd_p = torch.empty(1000, N)
    w["tree"].step_host(w["q"][:1000].cpu(), w["qd"][:1000].cpu(), qdd_p, goals=w["goals"][:1000].cpu(),
                        spheres=w["spheres"][:1000].cpu())
    assert torch.equal(qdd_p, w["qdd"][:1000].cpu())


def test_rollout_equals_stepwise_euler(world):
    """rmp2_rollout (control every 10 steps, dt = 0.01; the 100 Hz / 10 Hz loop of
    experiments/franka_panda/05_obstacle_avoidance.py:92-97 with an Euler integrator) equals the same
    loop written with single steps on the host side (consistency of the two entry points; the check against
    the oracle's trajectory is test_rollout_against_oracle_trajectory)."""
    w = world
    ns, fk, dev = w["ns"], w["fk"], w["dev"]
    # the damped full tree (config 5): a closed loop without a damping leaf is not a meaningful rollout
    core = S.build_config5(ns, fk, [0.5, 0.0, 0.5], N, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
    tree = core.compile(N, goal_leaves=["attractor"])
    Br, dt, n_steps, every = 4096, 0.01, 30, 10
    q, qd = w["q"][:Br].clone(), w["qd"][:Br].clone()
    goals, sph = w["goals"][:Br].contiguous(), w["spheres"][:Br].contiguous()
    qdd = torch.empty(Br, N, device=dev)
    tree.rollout(q, qd, qdd, dt, n_steps, every, goals=goals, spheres=sph)
    q2, qd2 = w["q"][:Br].clone(), w["qd"][:Br].clone()
    cmd = torch.empty(Br, N, device=dev)
    for step in range(n_steps):
        if step % every == 0:
            tree.step(q2, qd2, cmd, goals=goals, spheres=sph)
        qd2 = torch.addcmul(qd2, cmd, torch.tensor(dt, device=dev))
        q2 = torch.addcmul(q2, qd2, torch.tensor(dt, device=dev))
    assert torch.isfinite(q).all() and torch.isfinite(q2).all()
    # same arithmetic up to fma contraction of the integrator; chaotic amplification over 3 control
    # steps stays far below these bounds
    torch.testing.assert_close(q, q2, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(qd, qd2, rtol=1e-3, atol=1e-4)


def test_rollout_against_oracle_trajectory(world):
    """rmp2_rollout vs the ORACLE'S closed loop (oracle/harness.rollout, float64): 256 well-behaved scenes
    (tests/gpu_common.closed_loop_scene) of the full tree, 100 simulation steps, a control step every 10 (tests/golden/rollout_config5_n7.npz).  The float32 oracle
    trajectory is the yardstick for what float32 costs over a trajectory.  Error-growth bound: every control step
    contributes a command error of ~1e-6 |qdd| (the single-step parity bar), integrated twice over at most 1 s:
    |dq| <= 10 steps x 1e-5 x max|qdd| x T^2 / 2 -- with max|qdd| ~ 50 rad/s^2, 2.5e-3; the measured
    maxima are far below that."""
    import os
    from conftest import GOLDEN
    w = world
    ns, fk, dev = w["ns"], w["fk"], w["dev"]
    g = np.load(os.path.join(GOLDEN, "rollout_config5_n7.npz"))
    core = S.build_config5(ns, fk, [0.5, 0.0, 0.5], N, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
    tree = core.compile(N, goal_leaves=["attractor"])
    Br = g["q0"].shape[0]
    q, qd = torch.as_tensor(g["q0"], device=dev).clone(), torch.as_tensor(g["qd0"], device=dev).clone()
    goals = torch.as_tensor(g["goal"], device=dev).reshape(Br, 1, 3).contiguous()
    sph = torch.as_tensor(g["spheres"], device=dev)
    qdd = torch.empty(Br, N, device=dev)
    tree.rollout(q, qd, qdd, float(g["dt"]), int(g["n_steps"]), int(g["control_every"]), goals=goals, spheres=sph)
    q, qd = q.cpu().numpy().astype(np.float64), qd.cpu().numpy().astype(np.float64)
    err_q, err_qd = np.abs(q - g["q64"]).max(axis=1), np.abs(qd - g["qd64"]).max(axis=1)
    yard_q, yard_qd = np.abs(g["q32"] - g["q64"]).max(axis=1), np.abs(g["qd32"] - g["qd64"]).max(axis=1)
    stats = dict(median_err_q=float(np.median(err_q)), max_err_q=float(err_q.max()), median_yard_q=float(np.median(yard_q)),
                 max_yard_q=float(yard_q.max()), median_err_qd=float(np.median(err_qd)), max_err_qd=float(err_qd.max()),
                 median_yard_qd=float(np.median(yard_qd)), max_yard_qd=float(yard_qd.max()))
    print("rollout vs oracle:", stats)
    assert np.isfinite(q).all()
    assert err_q.max() <= 2.5e-3 and err_qd.max() <= 2.5e-2, stats
    assert np.median(err_q) <= 3 * np.median(yard_q) + 1e-7 and np.median(err_qd) <= 3 * np.median(yard_qd) + 1e-6, stats
    assert np.quantile(err_q, 0.99) <= 4 * np.quantile(yard_q, 0.99) + 1e-6, stats


def test_closed_loop_reaches_the_goal_without_penetration(world):
    """The reference's success criterion (experiments/franka_panda/06_cluttered_environment.py:120-131: control at
    10 Hz, simulate at 100 Hz until |x_ee - x_goal| < 0.02 m) on the GPU: the cluttered-environment tree of that
    script (TargetAttractor, JointVelocityCap, JointDamping, CSpaceBiasing, ObstacleAvoidance on every collision
    frame), from the ready pose, with sphere obstacles around and goals inside the arm's workspace.  Every
    environment must reach its goal and no collision-frame origin may ever be inside a sphere.

    This tree has no joint-limit leaf and its JointVelocityCap metric has poles at |qd| = max_velocity - 2 * damping
    region = 0.2 rad/s (rmp2.py:100-107, 1/0 in the reference as well): an environment that gets pinned there runs away
    and its command turns non-finite in float32 in the reference too (found on a B200: one of these 512 environments,
    qd = (.., -0.19999999, -0.20000005, ..), oracle metric entries of -3e5 / +8e4 in float64 and inf in float32).  Such
    environments are counted -- at most 1 % -- and left out of the clearance statistics."""
    w = world
    ns, fk, dev = w["ns"], w["fk"], w["dev"]
    from gpu_common import closed_loop_scene
    Bc, O_, dt, every = 512, 8, 0.01, 10
    q0, qd0, goal, sph = closed_loop_scene(Bc, O_, seed=7)
    q, qd = torch.as_tensor(q0, device=dev), torch.as_tensor(qd0, device=dev)
    frames = S.collision_frames(fk)
    ee = lambda: fk.forward(q, S.EE_FRAME)[:, :3, 3]
    origins = lambda: torch.stack([fk.forward(q, fr)[:, :3, 3] for fr in frames], dim=1)            # [B,K,3]
    spheres = torch.as_tensor(sph, device=dev)
    goals = torch.as_tensor(goal, device=dev).reshape(Bc, 1, 3).contiguous()
    core = S.build_config3(ns, fk, [0.5, 0.0, 0.5], N, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
    tree = core.compile(N, goal_leaves=["attractor"])
    qdd = torch.empty(Bc, N, device=dev)
    reached_at = torch.full((Bc,), -1.0, device=dev)
    clearance = torch.full((Bc,), 1e9, device=dev)
    t, horizon = 0.0, 60.0
    while t < horizon:
        tree.rollout(q, qd, qdd, dt, 50, every, goals=goals, spheres=spheres)                        # 0.5 s of simulated time
        t += 0.5
        alive = torch.isfinite(q).all(dim=1) & torch.isfinite(qd).all(dim=1)
        dist = torch.linalg.norm(ee() - goals[:, 0], dim=1)
        reached_at = torch.where((reached_at < 0) & alive & (dist < 0.02), torch.full_like(reached_at, t), reached_at)
        d = torch.linalg.norm(origins()[:, :, None, :] - spheres[:, None, :, :3], dim=-1) - spheres[:, None, :, 3]
        dmin = d.reshape(Bc, -1).min(dim=1).values
        clearance = torch.where(alive, torch.minimum(clearance, dmin), clearance)
        if bool(((reached_at >= 0) | ~alive).all()):
            break
    done = (reached_at >= 0).float().mean().item()
    lost = int((~alive).sum().item())
    print(f"closed loop: {100 * done:.1f} % of {Bc} environments within 0.02 m after {t:.1f} s simulated "
          f"(median {reached_at[reached_at >= 0].median().item():.1f} s), minimum clearance {clearance.min().item():.4f} m, "
          f"final |qd| max {qd[alive].abs().max().item():.3f}, non-finite (velocity-cap pole) {lost}")
    assert lost <= Bc // 100, f"{lost} environments turned non-finite"
    assert clearance.min().item() > 0.0, "a collision-frame origin entered a sphere"
    assert done >= 0.97, f"only {100 * done:.1f} % reached the goal"


def test_rollout_is_cuda_graph_capturable(world):
    """Every launch of rmp2_step / rmp2_rollout is asynchronous on the caller's stream and, after a first
    (warm-up) call has sized the tree's scratch, allocates nothing: a whole closed-loop rollout can be
    captured in a CUDA graph and replayed -- the way to run small, launch-bound batches.  The replay must
    reproduce the eager result bit for bit."""
    w = world
    ns, fk, dev = w["ns"], w["fk"], w["dev"]
    core = S.build_config5(ns, fk, [0.5, 0.0, 0.5], N, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
    tree = core.compile(N, goal_leaves=["attractor"])
    Br, dt, n_steps, every = 2048, 0.01, 30, 10
    goals, sph = w["goals"][:Br].contiguous(), w["spheres"][:Br].contiguous()
    q0, qd0 = w["q"][:Br].clone(), w["qd"][:Br].clone()
    q, qd, qdd = q0.clone(), qd0.clone(), torch.empty(Br, N, device=dev)
    tree.rollout(q, qd, qdd, dt, n_steps, every, goals=goals, spheres=sph)       # eager (also the warm-up)
    want_q, want_qd, want_qdd = q.clone(), qd.clone(), qdd.clone()
    gq, gqd, gqdd = q0.clone(), qd0.clone(), torch.empty(Br, N, device=dev)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            tree.rollout(gq, gqd, gqdd, dt, n_steps, every, goals=goals, spheres=sph)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(2):                                                              # replay twice from the same start
        gq.copy_(q0)
        gqd.copy_(qd0)
        graph.replay()
        torch.cuda.synchronize()
        for got, want in ((gq, want_q), (gqd, want_qd), (gqdd, want_qdd)):
            assert torch.equal(got.view(torch.int32), want.view(torch.int32))      # bit patterns
    # timing, informational: eager launches vs one graph launch
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for _ in range(5):
        tree.rollout(q, qd, qdd, dt, n_steps, every, goals=goals, spheres=sph)
    ev[1].record()
    ev[2].record()
    for _ in range(5):
        graph.replay()
    ev[3].record()
    torch.cuda.synchronize()
    print(f"rollout of {n_steps} sim steps ({n_steps // every} control steps), {Br} envs: eager "
          f"{ev[0].elapsed_time(ev[1]) / 5:.3f} ms, CUDA graph replay {ev[2].elapsed_time(ev[3]) / 5:.3f} ms")


def test_specialized_kernels_full_size(world):
    """The NVRTC-specialised frames / step kernels at the full benchmark size: bit-identical to the generic
    kernels on all 1,048,576 environments, with and without the exact early-out."""
    w = world
    core = S.build_config4(w["ns"], w["fk"], [0.5, 0.0, 0.5], N, lambda fr: w["ns"].TaskmapJointFrame4x4ToSphereDistance())
    tree = core.compile(N, goal_leaves=["attractor"])
    tree.specialize()
    out = torch.empty(B_FULL, N, device=w["dev"])
    for early in (True, False):
        tree.set_early_out(early)
        tree.step(w["q"], w["qd"], out, goals=w["goals"], spheres=w["spheres"])
        assert torch.equal(out.view(torch.int32), w["qdd"].view(torch.int32))
