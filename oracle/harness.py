"""CPU ORACLE harness -- TEST INFRASTRUCTURE ONLY (see oracle/rmp_oracle.py header).

Runs the restated reference on the workloads of ``riemannian_motion_policies_b200.scenarios``:
  * ``evaluate_loop``  one environment per ``RmpCore.evaluate`` call in a Python loop -- the way the
    reference itself runs (one env, eager autodiff per leaf); this is the timed CPU baseline;
  * ``evaluate_vmap``  the same single-environment function under ``torch.func.vmap`` -- identical
    arithmetic, used by the parity tests to check thousands of environments in seconds.
"""
import functools
import types

import numpy as np
import torch
from torch.func import vmap

from . import rmp_oracle as O
from riemannian_motion_policies_b200 import scenarios as S


def namespace(dtype=torch.float32):
    """The oracle's classes under the reference's names, bound to one dtype."""
    ns = types.SimpleNamespace(**{k: v for k, v in vars(O).items() if not k.startswith("_")})
    ns.RmpCore = functools.partial(O.RmpCore, dtype=dtype)
    ns.UrdfForwardKinematic = functools.partial(O.UrdfForwardKinematic, dtype=dtype)
    return ns


def make_fkine(n, dtype=torch.float32, robot=None):
    if robot == "gantry":
        return O.UrdfForwardKinematic(S.GANTRY_URDF, S.GANTRY_ORDER, dtype=dtype)
    if n == 2:
        return O.UrdfForwardKinematic(S.TWO_JOINT_URDF, S.TWO_JOINT_ORDER, dtype=dtype)
    if n == 7:
        return O.UrdfForwardKinematic(S.PANDA_WO_TOOL_URDF, S.PANDA_ORDER_7, dtype=dtype)
    if n == 9:
        return O.UrdfForwardKinematic(S.PANDA_URDF, S.PANDA_ORDER_9, dtype=dtype)
    raise ValueError(n)


def frame_origins(fkine, q, frames):
    """[K,3] origins of ``frames`` at configuration q [n] (what the simulator would report as
    pos_on_link for a control point at the frame origin)."""
    return torch.stack([fkine.forward(q[None], fr)[0, :3, 3] for fr in frames])


def _single_env_fn(config, n, fkine, dtype, combine=False):
    ns = namespace(dtype)
    frames = S.collision_frames(fkine)

    def one(q, qd, goal, spheres):
        q, qd, goal = q.to(dtype), qd.to(dtype), goal.to(dtype)
        if config == 1:
            core = S.build_config1(ns, fkine, goal)
        elif config == 2:
            core = S.build_config2(ns, fkine, goal, n)
        else:
            sph = spheres.to(dtype)
            origins = frame_origins(fkine, q, frames)                      # [K,3]
            r = origins[:, None, :] - sph[None, :, :3]
            on_obst = sph[None, :, :3] + sph[None, :, 3:4] * r / torch.linalg.norm(r, dim=-1, keepdim=True)
            on_link = origins[:, None, :].expand_as(on_obst)
            idx = {fr: i for i, fr in enumerate(frames)}
            tm_for = lambda fr: ns.TaskmapJointFrame4x4ToDistance(on_link[idx[fr]], on_obst[idx[fr]])
            core = S.BUILDERS[config](ns, fkine, goal, n, tm_for)
        return core.combine(q, qd) if combine else core.evaluate(q, qd)

    return one


def evaluate_loop(config, n, q, qd, goal, spheres=None, dtype=torch.float32, fkine=None):
    """Reference-style execution: one environment per call.  Inputs are numpy [B,...]."""
    fkine = fkine or make_fkine(n, dtype, robot="gantry" if config == 6 else None)
    one = _single_env_fn(config, n, fkine, dtype)
    out = []
    for b in range(q.shape[0]):
        sph = torch.as_tensor(spheres[b]) if spheres is not None else torch.zeros(0, 4)
        out.append(one(torch.as_tensor(q[b]), torch.as_tensor(qd[b]), torch.as_tensor(goal[b]), sph))
    return torch.stack(out).numpy()


def evaluate_vmap(config, n, q, qd, goal, spheres=None, dtype=torch.float32, chunk=1024, fkine=None):
    """Same arithmetic, vectorised over environments with torch.func.vmap."""
    fkine = fkine or make_fkine(n, dtype, robot="gantry" if config == 6 else None)
    one = _single_env_fn(config, n, fkine, dtype)
    B = q.shape[0]
    if spheres is None:
        spheres = np.zeros((B, 1, 4), dtype=np.float32)
    outs = []
    for s in range(0, B, chunk):
        sl = slice(s, min(B, s + chunk))
        outs.append(vmap(one)(torch.as_tensor(q[sl]), torch.as_tensor(qd[sl]), torch.as_tensor(goal[sl]),
                              torch.as_tensor(spheres[sl])))
    return torch.cat(outs).numpy()


def combined_vmap(config, n, q, qd, goal, spheres=None, dtype=torch.float64, chunk=1024, fkine=None):
    """(f [B,n], M [B,n,n]) before the resolve -- lets tests measure cond(M) and the singular-value
    gap around the pinv cutoff (SURVEY.md section 8c guards)."""
    fkine = fkine or make_fkine(n, dtype, robot="gantry" if config == 6 else None)
    one = _single_env_fn(config, n, fkine, dtype, combine=True)
    B = q.shape[0]
    if spheres is None:
        spheres = np.zeros((B, 1, 4), dtype=np.float32)
    fs, Ms = [], []
    for s in range(0, B, chunk):
        sl = slice(s, min(B, s + chunk))
        f, M = vmap(one)(torch.as_tensor(q[sl]), torch.as_tensor(qd[sl]), torch.as_tensor(goal[sl]),
                         torch.as_tensor(spheres[sl]))
        fs.append(f)
        Ms.append(M)
    return torch.cat(fs).numpy(), torch.cat(Ms).numpy()


def rollout(config, n, q, qd, goal, spheres, dt, n_steps, control_every, dtype=torch.float64, fkine=None):
    """Closed loop the way the experiments run it (experiments/franka_panda/05_obstacle_avoidance.py:92-97: control
    every `control_every` simulation steps, command held in between) with the simulator replaced by the explicit
    Euler integrator of rmp2_rollout:  qd += qdd dt;  q += qd dt  (velocity first).  Obstacles and goals stay fixed.
    Inputs numpy [B, ...]; returns (q, qd, last qdd) as numpy in `dtype`."""
    np_dtype = np.float64 if dtype == torch.float64 else np.float32
    q, qd = np.array(q, dtype=np_dtype), np.array(qd, dtype=np_dtype)
    dt = np_dtype(dt)
    qdd = np.zeros_like(q)
    for step in range(n_steps):
        if step % control_every == 0:
            qdd = evaluate_vmap(config, n, q, qd, goal, spheres, dtype=dtype, fkine=fkine).astype(np_dtype)
        qd = qd + qdd * dt
        q = q + qd * dt
    return q, qd, qdd


EPS32 = float(np.finfo(np.float32).eps)


def perturbed(arrays, k):
    """float64 copies of `arrays` [B, ...] with every component moved by a relative eps32 * U(-1, 1) (seeded by k):
    what merely ROUNDING THE INPUTS to float32 differently does.  The draw has the shape of ONE environment and is
    shared by the batch, so an environment's sensitivity does not depend on what else is in the batch.  None stays None."""
    rng = np.random.RandomState(1234 + k)
    return [None if a is None else np.asarray(a, dtype=np.float64) * (1.0 + EPS32 * rng.uniform(-1, 1, size=np.shape(a)[1:]))
            for a in arrays]


def sensitivity(fn, arrays, K=4):
    """Condition of a step for float32 inputs, per environment: the largest relative change of fn's float64 output
    over K seeded eps32-relative perturbations of all inputs.  fn(*arrays_f64) -> [B, n].  A float32 evaluation
    cannot be expected to be closer to the float64 truth than a small multiple of this (backward-error view: the
    result is the exact result for inputs that differ by rounding); it covers every amplification mechanism at once
    -- the conditioning of the combined metric, the 1/std_dev gains inside the obstacle leaf, the poles of the
    velocity-cap metric, a pinv truncation about to flip."""
    base = fn(*[None if a is None else np.asarray(a, dtype=np.float64) for a in arrays])
    S = np.zeros(base.shape[0])
    den = np.maximum(np.linalg.norm(base, axis=-1), 1e-30)
    for k in range(K):
        out = fn(*perturbed(arrays, k))
        S = np.maximum(S, np.linalg.norm(out - base, axis=-1) / den)
    return S


def metric_conditioning(M):
    """kappa * eps32 per environment, kappa = sigma_max / smallest singular value tf.linalg.pinv keeps (float32
    cutoff): how far pinv(M) f moves, relatively, when float32 accumulation leaves its unstructured eps32 * |M| noise
    on the combined metric.  Perturbing the INPUTS instead (`sensitivity`) keeps the structure of J^T A J -- an exactly
    rank-deficient sum stays rank deficient -- and so cannot see the conditioning of a metric whose small kept singular
    values sit just above the cutoff (config 4)."""
    s = np.linalg.svd(np.asarray(M, dtype=np.float64), compute_uv=False)
    n = s.shape[-1]
    cut = 10 * n * EPS32 * s[:, :1]
    kappa = s[:, 0] / np.where(s > cut, s, np.inf).min(-1)
    return kappa * EPS32


def config_sensitivity(config, n, q, qd, goal, spheres=None, K=4, fkine=None):
    """Float32 conditioning of the scenario trees per environment: the larger of `sensitivity` (inputs rounded
    differently) and `metric_conditioning` (float32 noise on the combined metric), float64 oracle under vmap."""
    fk = fkine or make_fkine(n, torch.float64, robot="gantry" if config == 6 else None)
    fn = lambda q_, qd_, goal_, sph_: evaluate_vmap(config, n, q_, qd_, goal_, sph_, dtype=torch.float64, fkine=fk)
    S_in = sensitivity(fn, [q, qd, goal, spheres], K=K)
    f64, M64 = combined_vmap(config, n, np.asarray(q, dtype=np.float64), np.asarray(qd, dtype=np.float64),
                             np.asarray(goal, dtype=np.float64),
                             None if spheres is None else np.asarray(spheres, dtype=np.float64), dtype=torch.float64, fkine=fk)
    return np.maximum(S_in, metric_conditioning(M64))
