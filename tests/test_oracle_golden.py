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
