"""Oracle trajectories for the closed-loop rollout test (CPU, a few minutes).

    python tests/golden/make_rollout_fixture.py      -> tests/golden/rollout_config5_n7.npz

256 seeded scenes (tests/gpu_common.closed_loop_scene: start near the ready pose at rest, goal in the workspace,
8 spheres around) of the full tree (config 5), 100 simulation steps of dt = 0.01 with a control step every 10 (the 100 Hz / 10 Hz loop of experiments/franka_panda/05_obstacle_avoidance.py:92-97 with the
simulator replaced by explicit Euler, see oracle/harness.rollout), in float64 (truth) and float32 (the yardstick
for what float32 arithmetic costs over a trajectory)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_common import closed_loop_scene                          # noqa: E402                                # noqa: E402
from oracle import harness as H                                    # noqa: E402

CONFIG, N, B, DT, STEPS, EVERY = 5, 7, 256, 0.01, 100, 10


def main():
    q, qd, goal, sph = closed_loop_scene(B, 8, seed=41)
    out = dict(q0=q, qd0=qd, goal=goal, spheres=sph, dt=DT, n_steps=STEPS, control_every=EVERY)
    for name, dtype in (("64", torch.float64), ("32", torch.float32)):
        qf, qdf, qddf = H.rollout(CONFIG, N, q, qd, goal, sph, DT, STEPS, EVERY, dtype=dtype)
        out["q" + name], out["qd" + name], out["qdd" + name] = qf, qdf, qddf
        print("float" + name, "done")
    np.savez_compressed(os.path.join(HERE, f"rollout_config{CONFIG}_n{N}.npz"), **out)


if __name__ == "__main__":
    main()
