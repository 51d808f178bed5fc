"""Multi-process path on CPU: world_size 2, gloo.  Covers the environment partition and the result
collection of riemannian_motion_policies_b200.sharding; the step function of each rank is the CPU
oracle (the CUDA kernels cannot run here), so the test also shows that sharding does not change results."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from riemannian_motion_policies_b200 import sharding


def test_shard_bounds_cover_the_batch():
    for B in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [sharding.shard_bounds(B, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import harness as H
        from riemannian_motion_policies_b200 import scenarios as S
        n = 7
        q, qd, goal = S.sample_panda_state(B, n, seed=5)

        def step_fn(ql, qdl, gl, sl):
            out = H.evaluate_vmap(2, n, ql.numpy(), qdl.numpy(), gl.numpy(), None)
            return torch.as_tensor(out)

        stepper = sharding.ShardedStep(step_fn)
        local = stepper.step_local(torch.as_tensor(q), torch.as_tensor(qd), torch.as_tensor(goal))
        lo, hi = sharding.shard_bounds(B, world, rank)
        assert local.shape == (hi - lo, n)
        full = stepper.step_and_gather(torch.as_tensor(q), torch.as_tensor(qd), torch.as_tensor(goal))
        assert full.shape == (B, n)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [11, 12])   # odd batch: ranks hold 6 and 5 environments (padded gather); even: one direct collective
def test_two_ranks_gloo_match_single_process(tmp_path, B):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    from oracle import harness as H
    from riemannian_motion_policies_b200 import scenarios as S
    q, qd, goal = S.sample_panda_state(B, 7, seed=5)
    ref = H.evaluate_vmap(2, 7, q, qd, goal, None)
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npy"))
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-7)
