"""GPU box helper: run the CUDA step on the golden fixtures and on seeded batches and save the raw
outputs to gpurun_out/ for offline error analysis in the build container."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gpu_common import make_inputs, product_evaluate                 # noqa: E402
from riemannian_motion_policies_b200 import scenarios as S           # noqa: E402

ns = S.product_namespace()
out = {}
for config, n in ((1, 2), (2, 7), (3, 7), (4, 7), (5, 7)):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"config{config}_n{n}.npz"))
    sph = g["spheres"] if "spheres" in g else None
    out[f"golden_config{config}_n{n}"] = product_evaluate(ns, config, n, g["q"], g["qd"], g["goal"], sph)
for config, n, B in ((4, 7, 1024), (5, 7, 1024), (3, 7, 2048)):
    q, qd, goal, sph = make_inputs(config, n, B)
    out[f"seeded_config{config}_n{n}_B{B}"] = product_evaluate(ns, config, n, q, qd, goal, sph)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "gpu_outputs.npz"), **out)
print("saved", list(out))
