"""B = 1 latency of core.evaluate(q, qd).numpy() (the reference's call) with a cProfile breakdown.
usage (GPU box): python tools/b1_latency_probe.py"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from riemannian_motion_policies_b200 import scenarios as S  # noqa: E402

ns = S.product_namespace()
n = 7
fk = ns.UrdfForwardKinematic(S.PANDA_WO_TOOL_URDF, S.PANDA_ORDER_7)
rng = np.random.RandomState(0)
q = rng.uniform(S.PANDA_Q_LOW[:n], S.PANDA_Q_HIGH[:n]).astype(np.float32)
qd = rng.uniform(-0.3, 0.3, size=n).astype(np.float32)
sph = S.sample_spheres(1, 64, 5)[0]
for name, core, kw in (("config2", S.build_config2(ns, fk, [0.5, 0.0, 0.5], n), {}),
                       ("config4_64_spheres", S.build_config4(ns, fk, [0.5, 0.0, 0.5], n, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance()), {"spheres": sph})):
    for _ in range(50):
        core.evaluate(q, qd, **kw).numpy()
    ts = []
    for _ in range(500):
        t0 = time.perf_counter()
        core.evaluate(q, qd, **kw).numpy()
        ts.append(time.perf_counter() - t0)
    print(name, "median us %.1f  p10 %.1f  p90 %.1f" % tuple(1e6 * np.quantile(ts, [0.5, 0.1, 0.9])))
    tree = core.compile(n)
    dev = torch.device("cuda")
    qt, qdt = torch.as_tensor(q, device=dev)[None], torch.as_tensor(qd, device=dev)[None]
    st = torch.as_tensor(sph, device=dev)[None] if kw else None
    qdd = torch.empty(1, n, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(10):
        tree.step(qt, qdt, qdd, spheres=st)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        tree.step(qt, qdt, qdd, spheres=st)
    e1.record()
    torch.cuda.synchronize()
    print(name, "device-resident step (launches only), us per step: %.1f" % (e0.elapsed_time(e1) * 10))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(500):
        core.evaluate(q, qd, **kw).numpy()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(12)
