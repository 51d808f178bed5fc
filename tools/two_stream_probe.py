"""Probe: does splitting a batch into chunks on two CUDA streams (kernels of different chunks co-resident on the SMs)
beat one stream?  Config 4, 2^20 environments.  Prints ms per full batch for 1 stream and for C chunks over 2 streams.
usage (GPU box): python tools/two_stream_probe.py [chunks ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                        # noqa: E402
from riemannian_motion_policies_b200 import scenarios as S          # noqa: E402

config, n, B = int(os.environ.get("PROBE_CONFIG", "4")), 7, 1 << 20
O_ = S.N_SPHERES[config]
dev = torch.device("cuda:0")
ns = S.product_namespace()
fk = ns.UrdfForwardKinematic(S.PANDA_WO_TOOL_URDF, S.PANDA_ORDER_7)
q, qd, goal, spheres = S.synth_inputs_device(fk, n, B, O_, 2, seed=5, device=dev)
goals = goal.reshape(B, 1, 3).contiguous()
qdd = torch.empty(B, n, device=dev)
early = bool(int(os.environ.get("PROBE_EARLY", "0")))
prio = bool(int(os.environ.get("PROBE_PRIO", "0")))


def timed(fn, reps=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


_, tree, _ = bench.build_tree(ns, S, fk, config, n)
tree.set_early_out(early)
tree.reserve(B, O_)
base = timed(lambda i: tree.step(q, qd, qdd, goals=goals, spheres=spheres[i % 2]))
print(f"one stream: {base:.4f} ms")
ref = qdd.clone()

for chunks in [int(a) for a in sys.argv[1:]] or [2, 4, 8]:
    per = B // chunks
    trees = []
    for s in range(2):
        _, t, _ = bench.build_tree(ns, S, fk, config, n)
        t.set_early_out(early)
        t.reserve(per, O_)
        trees.append(t)
    streams = [torch.cuda.Stream(priority=-1 if (prio and s == 1) else 0) for s in range(2)]
    main = torch.cuda.current_stream()

    def run(i):
        ev = torch.cuda.Event()
        ev.record(main)
        for s in streams:
            s.wait_event(ev)
        for c in range(chunks):
            sl = slice(c * per, (c + 1) * per)
            with torch.cuda.stream(streams[c % 2]):
                trees[c % 2].step(q[sl], qd[sl], qdd[sl], goals=goals[sl], spheres=spheres[i % 2][sl])
        for s in streams:
            e = torch.cuda.Event()
            e.record(s)
            main.wait_event(e)

    ms = timed(run)
    torch.cuda.synchronize()
    run(1)
    torch.cuda.synchronize()
    same = bool(torch.equal(qdd, ref))
    print(f"{chunks} chunks on 2 streams: {ms:.4f} ms ({base / ms:.3f}x), bit-identical to one stream: {same}")
