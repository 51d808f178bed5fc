"""Parity study on a B200 (GPU box): where the distance between the CUDA step and the oracle comes from.

    python tools/parity_study.py [--out gpurun_out/parity_study.json]

(1) Solver spread.  The same float32 (M, f) of 4096 config-4 environments (rank-deficient tree; oracle, float32
    accumulation) is solved by four independent float32 implementations of tf.linalg.pinv's rule -- LAPACK gesdd,
    LAPACK gesvd (SciPy), this library's direct (rank-revealing) solve + Jacobi fallback, and its Jacobi SVD
    alone -- and compared with the float64 solution of the SAME float32 matrix.  The error of every one of them
    scales with kappa * eps32 (kappa = sigma_max / smallest kept singular value); the study reports the
    quantiles of err / (kappa eps32), i.e. the constant a backward-stable float32 solve needs.
(2) Whole pipeline.  CUDA step vs oracle-f32 / oracle-f64 on the seeded 4096-environment batches of configs 4
    and 5 (tests/golden/parity_config*_n7.npz): how many environments pass each clause of the criterion
    (tests/gpu_common.assert_parity), and the distance from the float64 truth of the kernel and of the float32
    oracle itself in units of the environment's float32 conditioning S and of kappa * eps32.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

EPS32 = float(np.finfo(np.float32).eps)


def rel(a, b):
    return np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1e-30)


def tf_pinv_solve(M, f, dtype, svd):
    """x = pinv(M) f with TensorFlow's rule, SVD by `svd` in `dtype`."""
    n = M.shape[-1]
    rc = 10 * n * EPS32
    out = np.zeros(f.shape, dtype=np.float64)
    for i in range(M.shape[0]):
        U, s, Vt = svd(M[i].astype(dtype))
        keep = s > rc * s[0]
        out[i] = (Vt.T[:, keep] @ ((U.T[keep] @ f[i].astype(dtype)) / s[keep])).astype(np.float64)
    return out


def quantiles(x):
    x = np.asarray(x)
    return {k: float(np.quantile(x, q)) for k, q in (("q50", .5), ("q90", .9), ("q99", .99), ("q999", .999), ("max", 1.0))}


def solver_spread(path):
    import scipy.linalg as sl
    from riemannian_motion_policies_b200 import _native
    d = np.load(path)
    M32, f32 = d["M32"].astype(np.float32), d["f32"].astype(np.float32)
    B, n = f32.shape
    rc = 10 * n * EPS32
    truth = tf_pinv_solve(M32, f32, np.float64, lambda a: np.linalg.svd(a))        # exact solve of the SAME f32 problem
    s = np.linalg.svd(M32.astype(np.float64), compute_uv=False)
    cut = rc * s[:, :1]
    ratio = s / np.maximum(cut, 1e-300)
    near = ((ratio > 0.25) & (ratio < 4.0)).any(-1)
    kappa = s[:, 0] / np.where(s > cut, s, np.inf).min(-1)
    dev = torch.device("cuda")
    Md, fd = torch.as_tensor(M32, device=dev).contiguous(), torch.as_tensor(f32, device=dev).contiguous()
    sols = {}
    for name, mode in (("rmp2_direct+jacobi", 0), ("rmp2_jacobi_only", 1)):
        x = torch.empty(B, n, device=dev)
        _native.check(_native.lib().rmp2_pinv_solve(n, B, Md.data_ptr(), fd.data_ptr(), x.data_ptr(), 1, mode,
                                                    torch.cuda.current_stream().cuda_stream))
        sols[name] = x.cpu().numpy().astype(np.float64)
    sols["lapack_gesdd_f32"] = tf_pinv_solve(M32, f32, np.float32, lambda a: sl.svd(a, lapack_driver="gesdd"))
    sols["lapack_gesvd_f32"] = tf_pinv_solve(M32, f32, np.float32, lambda a: sl.svd(a, lapack_driver="gesvd"))
    keep = ~near
    out = {"envs": int(B), "near_cutoff_excluded": int(near.sum()), "median_kappa": float(np.median(kappa[keep])),
           "solvers": {}}
    for name, x in sols.items():
        e = rel(x, truth)[keep]
        out["solvers"][name] = {"err": quantiles(e), "err_over_kappa_eps32": quantiles(e / (kappa[keep] * EPS32)),
                                "frac_within_1e-5": float((e <= 1e-5).mean())}
    names = list(sols)
    out["pairwise_median"] = {f"{a} vs {b}": float(np.median(rel(sols[a], sols[b])[keep]))
                              for i, a in enumerate(names) for b in names[i + 1:]}
    return out


def pipeline(config, B=4096):
    from gpu_common import clause_counts, make_inputs, product_evaluate
    from riemannian_motion_policies_b200 import scenarios as S
    g = np.load(os.path.join(ROOT, "tests", "golden", f"parity_config{config}_n7.npz"))
    ns = S.product_namespace()
    q, qd, goal, sph = make_inputs(config, 7, B)
    got = product_evaluate(ns, config, 7, q, qd, goal, sph)
    np.save(os.path.join(ROOT, "gpurun_out", f"kernel_qdd_config{config}.npy"), got)      # for offline analysis
    return clause_counts(got, g["ref32"], g["ref64"], g["s64"], 7, sens=g["sens"] if "sens" in g else None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_study.json"))
    ap.add_argument("--mf", default=os.path.join(ROOT, "tests", "golden", "mf_config4_f32.npz"))
    args = ap.parse_args()
    res = {"lib": os.environ.get("RMP2_B200_LIB", "default build")}
    if os.path.exists(args.mf):
        res["solver_spread_config4"] = solver_spread(args.mf)
    for config in (4, 5):
        if os.path.exists(os.path.join(ROOT, "tests", "golden", f"parity_config{config}_n7.npz")):
            res[f"pipeline_config{config}"] = pipeline(config)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
