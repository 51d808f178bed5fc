"""Oracle outputs for the large seeded parity batches (build container or GPU box; CPU only, minutes).

    python tests/golden/make_parity_fixtures.py [config ...]      -> tests/golden/parity_config{c}_n7.npz

For B = 4096 seeded environments of configs 4 and 5 (inputs: tests/gpu_common.make_inputs, regenerated from
the seed by the tests) it stores what the parity criterion needs: the oracle's float32 and float64 outputs,
the singular values of the float64 combined metric (distance from the pinv cutoff), the float32-input sensitivity of
every environment (oracle/harness.sensitivity) and a checksum of the inputs.  Precomputed because the oracle needs minutes for 4096 environments of the 64-sphere trees -- time the
GPU tests and bench.py should not spend on the GPU box.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_common import make_inputs                                 # noqa: E402
from oracle import harness as H                                    # noqa: E402

B, N = 4096, 7


def input_digest(q, qd, goal, sph):
    h = hashlib.sha256()
    for a in (q, qd, goal, sph):
        if a is not None:
            h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    configs = [int(a) for a in sys.argv[1:]] or [4, 5]
    for config in configs:
        q, qd, goal, sph = make_inputs(config, N, B)
        ref32 = H.evaluate_vmap(config, N, q, qd, goal, sph, dtype=torch.float32)
        ref64 = H.evaluate_vmap(config, N, q, qd, goal, sph, dtype=torch.float64)
        _, M64 = H.combined_vmap(config, N, q, qd, goal, sph, dtype=torch.float64)
        s64 = np.linalg.svd(M64, compute_uv=False)
        sens = H.config_sensitivity(config, N, q, qd, goal, sph)      # float32-input condition of every environment
        out = os.path.join(HERE, f"parity_config{config}_n{N}.npz")
        np.savez_compressed(out, B=B, ref32=ref32.astype(np.float32), ref64=ref64, s64=s64, sens=sens,
                            digest=np.array(input_digest(q, qd, goal, sph)))
        print("wrote", out)
        if config == 4:       # the float32 combined (M, f) of the same batch: input of the solver study (tools/parity_study.py)
            f32, M32 = H.combined_vmap(config, N, q, qd, goal, sph, dtype=torch.float32)
            np.savez_compressed(os.path.join(HERE, "mf_config4_f32.npz"), M32=M32.astype(np.float32), f32=f32.astype(np.float32))


if __name__ == "__main__":
    main()
