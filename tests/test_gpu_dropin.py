"""CUDA step on the trees of the reference's experiment scripts, imported through compat/ (`from rmp import ...`),
against tests/golden/ref_exp_*.npz -- outputs of the scripts' own tree-building lines exec'd on top of the reference's
own modules (tests/golden/run_experiment_blocks_under_shim.py).  tests/test_dropin_experiments.py shows the trees built
here are descriptor-identical to what the verbatim script lines build on compat/."""
import os

import numpy as np
import pytest
import torch

import dropin_trees as DT
from conftest import GOLDEN
from gpu_common import assert_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key", sorted(DT.ROBOT))
def test_experiment_tree_matches_reference_run(key, native_lib):
    g = np.load(os.path.join(GOLDEN, f"ref_exp_{key}.npz"))
    ns = DT.compat_namespace()
    fk = DT.make_fkine(ns, key)
    dm = ns.Datamanager(fk)
    core = DT.build(ns, key, fk, dm)
    frames = list(g["frames"])
    got, ref64 = [], []
    for b in range(g["q"].shape[0]):
        lo, hi = int(g["row_count"][:b].sum()), int(g["row_count"][:b + 1].sum())
        if hi > lo:
            dm.update(g["q"][b], DT.unpack_rows(g["rows"][lo:hi], frames))       # the reference's feed, data_management.py:22-37
        out = core.evaluate(g["q"][b], g["qd"][b])                             # the reference's call: numpy in, .numpy() out
        got.append(out.numpy())
        ref64.append(DT.oracle_evaluate(key, g, b, torch.float64))
    stats = assert_parity(np.stack(got), g["qdd_ref"], np.stack(ref64), label=f"experiment block {key}")
    print(key, str(g["script"]), tuple(g["lines"]), stats)
