"""Reference vectors from the experiment scripts themselves (build container only; needs /root/reference).

    python tests/golden/run_experiment_blocks_under_shim.py      -> tests/golden/ref_exp_<key>.npz

For every script in tests/dropin.SCRIPTS the tree-building block is exec'd VERBATIM on top of the reference's own
`rmp.py / rmp2.py / kinematics.py / taskmap.py / data_management.py` (TensorFlow replaced by oracle/tf_shim,
PyBullet by a joint-table stub -- see tests/dropin.py), then `data_manager.update(q, distance_data)` and
`core.evaluate(q, qd)` run on seeded states and synthetic closest-point tuples in the reference's wire format
(simulation.py:462-484).  tests/test_dropin_experiments.py execs the SAME lines on top of this repo's compat/
modules and tests/test_gpu_dropin.py compares the CUDA step with these vectors.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dropin                                                       # noqa: E402

B = {"exp06": 24, "exp04": 16, "two01": 16, "two03": 16, "two05": 16}


def sample_state(key, b, rng):
    """Seeded q, qd (float32) for script `key` -- the robots' limits as in simulation.py:84-86,137-139."""
    if dropin.SCRIPTS[key][1] == "panda":
        idx = [0, 1, 2, 3, 4, 5, 6, 9, 10]
        lo, hi = dropin.PANDA_LIMITS[0][idx], dropin.PANDA_LIMITS[1][idx]
    else:
        lo, hi = -np.pi * np.ones(2), np.pi * np.ones(2)
        if key == "two03":                      # joint-limit leaf: stay inside, some environments close to a limit
            lo, hi = 0.999 * lo, 0.999 * hi
    q = rng.uniform(lo, hi).astype(np.float32)
    qd = rng.uniform(-0.4, 0.4, size=q.shape).astype(np.float32)
    return q, qd


def synth_distance_data(frames, origins, rotations, rng, k_max, d_lo, d_hi):
    """Closest-point tuples (frame, pos_on_link, pos_on_obstacle, normal, distance, description): link points within
    a few cm of the frame, obstacle points at distance d along a random direction, normal from obstacle to link."""
    data = []
    for fi, frame in enumerate(frames):
        for _ in range(rng.randint(0, k_max + 1)):
            on_link = origins[fi] + rotations[fi] @ rng.uniform(-0.06, 0.06, size=3)
            direction = rng.normal(size=3)
            direction /= np.linalg.norm(direction)
            dist = rng.uniform(d_lo, d_hi)
            data.append((frame, on_link.astype(np.float32), (on_link - dist * direction).astype(np.float32),
                         direction.astype(np.float32), np.float32(dist), "synthetic"))
    return data


def pack(data, frames):
    """distance_data -> float32 rows [K, 11] = (frame index, pos_on_link, pos_on_obstacle, normal, distance)."""
    rows = [np.concatenate([[frames.index(d[0])], d[1], d[2], d[3], [d[4]]]) for d in data]
    return np.array(rows, dtype=np.float32).reshape(-1, 11)


def unpack(rows, frames):
    return [(frames[int(r[0])], r[1:4].copy(), r[4:7].copy(), r[7:10].copy(), np.float32(r[10]), "synthetic") for r in rows]


def run(key):
    env = dropin.build_tree(key, "reference")
    tf = env["tf"]                                                  # the shim, imported by the script's own header
    core, fkine = env["core"], env["fkine"]
    frames = list(fkine.frame_names)
    rng = np.random.RandomState(1000 + sorted(dropin.SCRIPTS).index(key))
    out = dict(q=[], qd=[], qdd_ref=[], rows=[], row_count=[])
    feeds = key in ("exp06", "two05")
    import builtins
    real_print, builtins.print = builtins.print, (lambda *a, **k: None)
    try:
        for b in range(B[key]):
            q, qd = sample_state(key, b, rng)
            rows = np.zeros((0, 11), np.float32)
            if feeds:
                Ts = [fkine.forward(tf.constant([q], dtype=tf.float32), tf.constant(fr)).numpy()[0] for fr in frames]
                use = [i for i, fr in enumerate(frames) if key == "two05" or ("collision_avoidance_for_" + fr) in core.rmps]
                data = synth_distance_data([frames[i] for i in use], [Ts[i][:3, 3] for i in use], [Ts[i][:3, :3] for i in use],
                                           rng, k_max=4, d_lo=0.05, d_hi=(1.3 if key == "two05" else 0.6))
                if key == "two05":              # every leaf needs rows here: the v1 leaf has no empty-set semantics
                    have = {d[0] for d in data}
                    for i in use:
                        if frames[i] not in have:
                            data += synth_distance_data([frames[i]], [Ts[i][:3, 3]], [Ts[i][:3, :3]], rng, 1, 0.05, 1.3) or \
                                    synth_distance_data([frames[i]], [Ts[i][:3, 3]], [Ts[i][:3, :3]], np.random.RandomState(b), 1, 0.05, 1.3)
                rows = pack(data, frames)
                env["data_manager"].update(q, unpack(rows, frames))
            res = core.evaluate(q, qd).numpy()
            out["q"].append(q); out["qd"].append(qd); out["qdd_ref"].append(res)
            out["rows"].append(rows); out["row_count"].append(len(rows))
    finally:
        builtins.print = real_print
    d = dict(q=np.stack(out["q"]), qd=np.stack(out["qd"]), qdd_ref=np.stack(out["qdd_ref"]).astype(np.float32),
             rows=np.concatenate(out["rows"]), row_count=np.array(out["row_count"]), frames=np.array(frames),
             lines=np.array(env["_lines"]), script=np.array(dropin.SCRIPTS[key][0]))
    if "goal" in env:
        d["goal"] = np.array(env["goal"].base_position, dtype=np.float64)
    return d


def main():
    for key in (sys.argv[1:] or sorted(dropin.SCRIPTS)):
        d = run(key)
        np.savez_compressed(os.path.join(HERE, f"ref_exp_{key}.npz"), **d)
        print("wrote ref_exp_%s.npz" % key, d["qdd_ref"].shape, "rows", int(d["row_count"].sum()))


if __name__ == "__main__":
    main()
