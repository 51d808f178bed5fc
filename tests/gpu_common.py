"""Shared helpers of the -m gpu parity tests: product-side builders and the parity criterion."""
import numpy as np
import torch

from oracle import harness as H
from riemannian_motion_policies_b200 import scenarios as S

REL_TOL = 1e-5      # BASELINE.json north_star: 1e-5 relative on qddot in fp32


def product_fkine(ns, n):
    if n == 2:
        return ns.UrdfForwardKinematic(S.TWO_JOINT_URDF, S.TWO_JOINT_ORDER)
    if n == 7:
        return ns.UrdfForwardKinematic(S.PANDA_WO_TOOL_URDF, S.PANDA_ORDER_7)
    return ns.UrdfForwardKinematic(S.PANDA_URDF, S.PANDA_ORDER_9)


def product_core(ns, config, n, fkine):
    goal0 = [0.5, 0.0, 0.5]
    if config == 1:
        return S.build_config1(ns, fkine, goal0)
    if config == 2:
        return S.build_config2(ns, fkine, goal0, n)
    return S.BUILDERS[config](ns, fkine, goal0, n, lambda frame: ns.TaskmapJointFrame4x4ToSphereDistance())


def product_evaluate(ns, config, n, q, qd, goal, spheres=None, fkine=None, core=None):
    fkine = fkine or product_fkine(ns, n)
    core = core or product_core(ns, config, n, fkine)
    dev = torch.device("cuda")
    out = core.evaluate(torch.as_tensor(q, device=dev), torch.as_tensor(qd, device=dev),
                        goals=torch.as_tensor(goal, device=dev),
                        spheres=None if spheres is None else torch.as_tensor(spheres, device=dev))
    return out.cpu().numpy()


def rel_err(a, b):
    return np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1e-30)


def assert_parity(got, ref32, ref64, M64=None, n=None, label="", max_excluded=0.05):
    """Parity criterion (made explicit).  Per environment, with
    e32 = |got - ref32|/|ref32|, e64 = |got - ref64|/|ref64|, yard = |ref32 - ref64|/|ref64| (the float32
    restatement's own distance from the float64 truth) and kappa = sigma_max / smallest kept singular
    value of the combined metric M:

        e32 <= 1e-5                                   the north-star bar, or
        e64 <= max(1e-5, 2 * yard)                    not worse than the float32 reference itself, or
        e64 <= 64 * kappa * eps32                     backward-stable float32 bound for this env's M

    plus, over the batch, the kernel must be statistically as close to the truth as the float32
    restatement is: median(e64) <= 2 median(yard), q99(e64) <= 3 q99(yard).  Environments with a singular
    value within a factor 4 of the pinv cutoff are excluded (the truncation is discontinuous there);
    they must stay below 5 % of the batch (10 % for the rank-deficient config 4 tree, whose weak
    obstacle metrics put a continuum of singular values around the cutoff)."""
    eps32 = np.finfo(np.float32).eps
    e32 = rel_err(got, ref32)
    e64 = rel_err(got, ref64)
    yard = rel_err(ref32, ref64)
    excluded = np.zeros(e32.shape, dtype=bool)
    kappa = np.ones(e32.shape)
    if M64 is not None:
        s = np.linalg.svd(M64, compute_uv=False)
        cut = 10 * n * eps32 * s[:, :1]
        ratio = s / np.maximum(cut, 1e-300)
        excluded = ((ratio > 0.25) & (ratio < 4.0)).any(-1)
        kept = np.where(s > cut, s, np.inf)
        kappa = s[:, 0] / kept.min(-1)
    ok = (e32 <= REL_TOL) | (e64 <= np.maximum(REL_TOL, 2 * yard)) | (e64 <= 64 * kappa * eps32)
    bad = ~ok & ~excluded
    keep = ~excluded
    stats = dict(envs=int(len(e32)), frac_strict=float((e32[keep] <= REL_TOL).mean()), median_e32=float(np.median(e32[keep])),
                 median_e64=float(np.median(e64[keep])), median_yard=float(np.median(yard[keep])),
                 q99_e64=float(np.quantile(e64[keep], 0.99)), q99_yard=float(np.quantile(yard[keep], 0.99)),
                 median_kappa=float(np.median(kappa[keep])), excluded=int(excluded.sum()))
    assert not bad.any(), (f"{label}: {int(bad.sum())}/{len(bad)} envs out of tolerance; worst e32={e32[bad].max():.3e} "
                           f"e64={e64[bad].max():.3e} yard={yard[bad].max():.3e} kappa={kappa[bad].max():.3e} {stats}")
    assert excluded.mean() < max_excluded, f"{label}: too many envs near the pinv cutoff ({excluded.mean():.3f})"
    if keep.sum() >= 200:
        assert stats["median_e64"] <= 2 * stats["median_yard"] + 1e-7, f"{label}: {stats}"
        assert stats["q99_e64"] <= 3 * stats["q99_yard"] + 1e-6, f"{label}: {stats}"
    return stats


def make_inputs(config, n, B, seed=None):
    seed = S.SEEDS[config] if seed is None else seed
    if config == 1:
        q, qd, goal = S.sample_two_joint(B, seed)
        return q, qd, goal, None
    q, qd, goal = S.sample_panda_state(B, n, seed)
    O_ = S.N_SPHERES[config]
    if not O_:
        return q, qd, goal, None
    fk = H.make_fkine(n, torch.float64)
    frames = S.collision_frames(fk)
    origins = torch.func.vmap(lambda qq: H.frame_origins(fk, qq, frames))(torch.as_tensor(q).double()).numpy()
    return q, qd, goal, S.sample_spheres(B, O_, seed, origins)
