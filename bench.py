#!/usr/bin/env python
"""Benchmark of the RMP2 control-step hot path (contract: see the task brief / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2|3|4|5] [--envs B]

A "step" = one pass of the hot path (RmpCore.evaluate, reference rmp.py:133-155) over one batch of
synthetic environments.  Default workload: BASELINE.json configs[3] -- Franka Panda (7 DOF), target +
joint-limit + 64 dynamic sphere obstacles, 1,048,576 environments per GPU (weak scaling over N GPUs,
no collective on the step path).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "franka_rmp2_control_steps_per_sec"
UNIT = "env-steps/s"
# algorithmic work per environment-step, SURVEY.md section 8d (n = 7, panda_wo_tool)
FLOPS_PER_ENV = {2: 2704.0, 3: 16222.0, 4: 45147.0, 5: 45406.0}
BYTES_PER_ENV = {2: 96.0, 3: 352.0, 4: 1120.0, 5: 1120.0}
WORKLOAD = {
    2: "franka_panda_7dof_target+cspace_bias (BASELINE configs[1])",
    3: "franka_panda_7dof_cluttered_16_spheres (BASELINE configs[2])",
    4: "franka_panda_7dof_target+joint_limits+64_dynamic_spheres (BASELINE configs[3])",
    5: "franka_panda_7dof_full_tree_64_spheres (BASELINE configs[4])",
}
DEFAULT_ENVS = {2: 1 << 20, 3: 1 << 20, 4: 1 << 20, 5: 1 << 20}
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # analytic, at clocks.max.sm
# The pair loop's real bound (profiles/r1_pipe_peaks.json): FP32 and ALU-pipe instructions
# share lane time on a B200 SM (128 lanes/clk), MUFU (16/clk/SM) runs underneath.  Thread-level operations
# per (frame, sphere) pair, counted from the SASS of rmp2_spheres_kernel's unrolled loop:
LANE_OPS_PER_PAIR = 42.5        # 41 FP32 (35 packed halves + 6 scalar) + 1.5 ALU (copysign LOP3, loop overhead); round 1: 49
MUFU_PER_PAIR = 5.0
LANE_PEAK_TOPS = 148 * 128 * 1.965e9 / 1e12
MUFU_PEAK_TOPS = 148 * 16 * 1.965e9 / 1e12


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock, power and clock-event (throttle) reasons DURING the timed region: an NVML
    polling thread (every ~2 ms; the timed region can be shorter than nvidia-smi's fastest period),
    with an `nvidia-smi -lms` subprocess as fallback."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.samples = []          # (sm_mhz, sm_max_mhz, power_w, reasons bitmask)
        self.stop_flag = threading.Event()
        self.nvml = None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; map through the UUID-free common case
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                              "--format=csv,noheader,nounits", "-lms", "100"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._read, daemon=True)
                self.thread.start()
            except OSError:
                self.proc = None
        return self

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((float(sm), float(mx), float(pw), int(rs)))
            except Exception:
                break
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def __exit__(self, *exc):
        self.stop_flag.set()
        if self.nvml is not None and self.thread is not None:
            self.thread.join(timeout=1)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, power, reasons = [], [], [], set()
        if self.samples:
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            for a, b, c, r in self.samples:
                sm.append(a)
                mx.append(b)
                power.append(c)
                for name, bit in bits.items():
                    if r & bit:
                        reasons.add(name)
            source = "nvml"
        else:
            source = "nvidia-smi"
            for row in self.rows:
                parts = [p.strip() for p in row.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                    power.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(self.NAMES, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "reasons": sorted(reasons), "samples": len(sm), "source": source}


# --------------------------------------------------------------------------------------- CPU arm
_CPU = {}


def _cpu_init(config, n, n_spheres):
    """Pool initializer: import the oracle, build the kinematics once and pay torch's first-call cost."""
    import torch
    torch.set_num_threads(1)
    from oracle import harness as H
    from riemannian_motion_policies_b200 import scenarios as S
    _CPU.update(H=H, S=S, config=config, n=n, n_spheres=n_spheres, fk=H.make_fkine(n))
    _cpu_worker((999, 1))


def _cpu_worker(args):
    """One host core: restated reference (oracle), one environment per call like the reference."""
    seed, n_envs = args
    H, S, config, n, O_ = _CPU["H"], _CPU["S"], _CPU["config"], _CPU["n"], _CPU["n_spheres"]
    q, qd, goal = S.sample_panda_state(n_envs, n, seed)
    sph = S.sample_spheres(n_envs, O_, seed) if O_ else None
    t0 = time.perf_counter()
    H.evaluate_loop(config, n, q, qd, goal, sph, fkine=_CPU["fk"])
    return n_envs, time.perf_counter() - t0


class CpuReference:
    """The restated reference on `cores` host processes (1 thread each), warmed up once."""

    def __init__(self, config, n, cores, n_spheres):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_cpu_init, initargs=(config, n, n_spheres))
        self.step_id = 0

    def step(self, envs_per_core):
        """One bounded sample: every core evaluates `envs_per_core` environments; -> (env-steps/s, envs, seconds)."""
        self.step_id += 1
        jobs = [(1000 * self.step_id + i, envs_per_core) for i in range(self.cores)]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        total = sum(r[0] for r in res)
        return total / wall, total, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_reference_throughput(config, n, cores, envs_per_core, n_spheres, budget_s=20.0):
    """Bounded CPU sample: two warm-up rounds (Pool() returns before the workers have imported torch, so the
    first maps still pay for it), then up to five measured rounds within `budget_s`; -> (median env-steps/s,
    environments measured, seconds measured)."""
    ref = CpuReference(config, n, cores, n_spheres)
    try:
        for _ in range(2):
            ref.pool.map(_cpu_worker, [(7 + i, 1) for i in range(2 * cores)], chunksize=1)
        values, total, wall, t0 = [], 0, 0.0, time.perf_counter()
        for _ in range(5):
            v, envs, w = ref.step(envs_per_core)
            values.append(v)
            total += envs
            wall += w
            if time.perf_counter() - t0 > budget_s:
                break
        return float(np.median(values)), total, wall
    finally:
        ref.close()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path.  The reference is
    TensorFlow 2.10 + PyBullet and cannot be installed here (no wheels, no network), so the arm times
    the oracle port (oracle/rmp_oracle.py: same per-leaf autodiff structure, one env per call) on all
    host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    config, n = args.config, 7
    cores = min(host_cores(), 64)
    per_core = max(2, args.envs_per_core)
    O_ = {2: 0, 3: 16, 4: 64, 5: 64}[config]
    ref = CpuReference(config, n, cores, O_)
    values, t_start = [], time.perf_counter()
    try:
        for _ in range(max(0, args.warmup)):
            ref.step(1)
        for _ in range(max(1, args.steps)):
            values.append(ref.step(per_core)[0])
            if time.perf_counter() - t_start > args.reference_budget_s:     # bounded: a few minutes at most
                break
    finally:
        ref.close()
    value = float(np.median(values))
    sample = f"{cores} processes x {per_core} envs per step, one env per RmpCore.evaluate call (autodiff per leaf)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "steps_measured": len(values),
        "warmup": args.warmup, "ms_per_step": 1e3 * cores * per_core / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD[config], "envs_per_step": cores * per_core, "spheres_per_env": O_, "dof": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- GPU arm
def build_tree(ns, S, fk, config, n, specialize=True):
    """-> (core, tree, specialisation info) of one BASELINE config on the 7-DOF Panda."""
    goal0 = [0.5, 0.0, 0.5]
    sphere_tm = lambda frame: ns.TaskmapJointFrame4x4ToSphereDistance()
    core = S.build_config2(ns, fk, goal0, n) if config == 2 else S.BUILDERS[config](ns, fk, goal0, n, sphere_tm)
    tree = core.compile(n, goal_leaves=["target" if config == 2 else "attractor"])
    info = {"on": False, "nvrtc_seconds": None}
    if specialize:
        t_spec = time.perf_counter()
        try:
            info = {"on": True, "nvrtc_seconds": tree.specialize(), "wall_seconds": time.perf_counter() - t_spec}
        except (RuntimeError, NotImplementedError) as exc:   # no NVRTC on this box: the generic CUDA kernels run instead
            info = {"on": False, "nvrtc_seconds": None, "error": str(exc)[:200]}
    return core, tree, info


def fixture_parity(ns, config, n=7):
    """Parity on the seeded 4096-environment batch whose oracle outputs are committed
    (tests/golden/parity_config*_n7.npz): pass counts per clause of the criterion (tests/gpu_common.py)."""
    path = os.path.join(ROOT, "tests", "golden", f"parity_config{config}_n{n}.npz")
    if not os.path.exists(path):
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gpu_common import clause_counts, make_inputs, product_evaluate
    g = np.load(path)
    B = int(g["B"])
    q, qd, goal, sph = make_inputs(config, n, B)
    got = product_evaluate(ns, config, n, q, qd, goal, sph)
    out = clause_counts(got, g["ref32"], g["ref64"], g["s64"], n, sens=g["sens"] if "sens" in g else None)
    out["inputs"] = "tests/gpu_common.make_inputs seed, oracle outputs from tests/golden/" + os.path.basename(path)
    return out


def time_steps(fn, steps, torch, warmup=3):
    """ms per call of fn(i) over `steps` calls, CUDA events on the current stream."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def graph_of(fn, torch):
    """Capture fn() (already warmed up: scratch sized) into a CUDA graph -- the way to run launch-bound batches."""
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    return graph


def other_config_lines(ns, S, fk, n, device, steps, torch):
    """The other BASELINE configs at their own batch sizes on this GPU (N = 1 only): config 5 at 2^20 environments,
    configs 2 / 3 at 4,096 / 65,536 environments replayed from a CUDA graph (launch-bound sizes)."""
    out = {}
    for config, B, graphed in ((5, 1 << 20, False), (2, 4096, True), (3, 65536, True)):
        core, tree, spec = build_tree(ns, S, fk, config, n)
        O_ = S.N_SPHERES[config]
        tree.set_early_out(False)
        tree.set_merge_coincident(False)
        nb = 2 if O_ else 1
        q, qd, goal, spheres = S.synth_inputs_device(fk, n, B, O_, nb, seed=S.SEEDS[config], device=device)
        goals = goal.reshape(B, 1, 3).contiguous()
        qdd = torch.empty(B, n, device=device)
        step = lambda i: tree.step(q, qd, qdd, goals=goals, spheres=spheres[i % nb] if O_ else None)
        line = {"workload": WORKLOAD[config], "envs": B, "spheres_per_env": O_, "specialized": spec["on"]}
        if graphed:
            step(0)
            torch.cuda.synchronize()
            graph = graph_of(lambda: step(0), torch)
            ms = time_steps(lambda i: graph.replay(), steps, torch)
            line["launch"] = "CUDA graph replay"
            line["ms_per_step_eager"] = time_steps(step, steps, torch)
        else:
            ms = time_steps(step, steps, torch)
            line["launch"] = "eager"
        line["ms_per_step"] = ms
        line["value"] = B / (ms * 1e-3)
        line["unit"] = UNIT
        if O_:                                  # the exact early-out alone, then the library default; same launch mode
            for key, merged in (("value_early_out", False), ("value_library_default", True)):
                tree.set_early_out(True)
                tree.set_merge_coincident(merged)
                if graphed:
                    step(0)
                    torch.cuda.synchronize()
                    graph_eo = graph_of(lambda: step(0), torch)
                    ms_eo = time_steps(lambda i: graph_eo.replay(), steps, torch)
                else:
                    ms_eo = time_steps(step, steps, torch)
                line[key] = B / (ms_eo * 1e-3)
            line["obstacle_leaves_and_pair_loops"] = list(tree.obstacle_slots())
        line["parity"] = fixture_parity(ns, config, n)
        out[f"config{config}"] = line
        del tree, core, q, qd, goal, spheres, qdd
        torch.cuda.empty_cache()
    return out


def rollout_lines(ns, S, fk, n, device, torch):
    """Closed-loop rollout throughput (rmp2_rollout: 100 simulation steps of dt = 0.01, a control step every 10; full
    tree = config 5, 16 fixed spheres around a start near the ready pose -- scenarios.closed_loop_scene_device),
    replayed from a CUDA graph."""
    out = {}
    core, tree, spec = build_tree(ns, S, fk, 5, n)
    for B in (4096, 1 << 20):
        q0, qd0, goal, sph = S.closed_loop_scene_device(fk, n, B, 16, seed=11, device=device)
        spheres = [sph]
        goals = goal.reshape(B, 1, 3).contiguous()
        q, qd, qdd = q0.clone(), qd0.clone(), torch.empty(B, n, device=device)
        sim_steps, every = 100, 10
        run = lambda: tree.rollout(q, qd, qdd, 0.01, sim_steps, every, goals=goals, spheres=spheres[0])
        run()
        torch.cuda.synchronize()
        graph = graph_of(run, torch)
        reps = 5 if B > 100000 else 20

        def replay(i):
            q.copy_(q0)
            qd.copy_(qd0)
            graph.replay()

        ms = time_steps(replay, reps, torch, warmup=2)
        out[f"envs_{B}"] = {"ms_per_rollout": ms, "sim_steps": sim_steps, "control_steps": sim_steps // every,
                            "control_steps_per_s": B * (sim_steps // every) / (ms * 1e-3),
                            "sim_steps_per_s": B * sim_steps / (ms * 1e-3), "launch": "CUDA graph replay",
                            "finite": bool(torch.isfinite(q).all().item())}
        out[f"envs_{B}"]["spheres_per_env"] = 16
        del q0, qd0, goal, sph, spheres, goals, q, qd, qdd, graph
        torch.cuda.empty_cache()
    return out


def b1_latency(ns, S, fk, n, torch):
    """core.evaluate(q, qd) the way the reference's experiments call it: one environment, NumPy in, .numpy() out
    (experiments/franka_panda/05_obstacle_avoidance.py:96).  Median wall-clock microseconds over 200 calls."""
    out = {}
    rng = np.random.RandomState(0)
    q = rng.uniform(S.PANDA_Q_LOW[:n], S.PANDA_Q_HIGH[:n]).astype(np.float32)
    qd = rng.uniform(-0.3, 0.3, size=n).astype(np.float32)
    sph = S.sample_spheres(1, 64, 5)[0]
    for name, config, kw in (("config2", 2, {}), ("config4_64_spheres", 4, {"spheres": sph})):
        core, tree, _ = build_tree(ns, S, fk, config, n, specialize=False)
        core = S.build_config2(ns, fk, [0.5, 0.0, 0.5], n) if config == 2 else \
            S.build_config4(ns, fk, [0.5, 0.0, 0.5], n, lambda fr: ns.TaskmapJointFrame4x4ToSphereDistance())
        for _ in range(20):
            core.evaluate(q, qd, **kw).numpy()
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            core.evaluate(q, qd, **kw).numpy()
            ts.append(time.perf_counter() - t0)
        out[name] = float(np.median(ts) * 1e6)
    return out


def h2d_peak_gbs(torch, device, barrier, nbytes=1 << 30, reps=5):
    """Pinned host -> device copy rate of this rank, all ranks copying at the same time (plain cudaMemcpyAsync of
    one 1 GiB buffer per copy)."""
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dev.copy_(host, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del host, dev
    return gbs


def cpu_vectorised_baseline(config, n, cores, envs=1024):
    """BASELINE.md section 4 plan C: the same oracle arithmetic vectorised over environments (torch.func.vmap) with
    all host threads -- the strongest CPU comparator available without TensorFlow."""
    import torch
    from oracle import harness as H
    from riemannian_motion_policies_b200 import scenarios as S
    old = torch.get_num_threads()
    torch.set_num_threads(cores)
    try:
        O_ = S.N_SPHERES[config]
        q, qd, goal = S.sample_panda_state(envs, n, 123)
        sph = S.sample_spheres(envs, O_, 123) if O_ else None
        H.evaluate_vmap(config, n, q[:64], qd[:64], goal[:64], None if sph is None else sph[:64])      # warm-up
        t0 = time.perf_counter()
        H.evaluate_vmap(config, n, q, qd, goal, sph, chunk=256)
        dt = time.perf_counter() - t0
    finally:
        torch.set_num_threads(old)
    return {"value": envs / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{envs} envs of the same workload, oracle arithmetic under torch.func.vmap in chunks of 256, "
                      f"torch.set_num_threads({cores}), {dt:.1f} s"}


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from riemannian_motion_policies_b200 import _native, scenarios as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa_cores = 0
    if world > 1:
        # one process per GPU: keep this rank's pinned staging memory and its copy threads on the GPU's own socket
        from riemannian_motion_policies_b200.sharding import bind_host_to_gpu
        numa_cores = bind_host_to_gpu(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout on first use; keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    config, n = args.config, 7
    B = args.envs or DEFAULT_ENVS[config]
    O_ = S.N_SPHERES[config]
    ns = S.product_namespace()
    fk = ns.UrdfForwardKinematic(S.PANDA_WO_TOOL_URDF, S.PANDA_ORDER_7)
    # the product's path for large batches: frames / step kernels rebuilt for this tree by NVRTC (one-off,
    # outside the timed region like any warm-up); --no-specialize times the generic table-driven kernels
    core, tree, specialized = build_tree(ns, S, fk, config, n, specialize=not args.no_specialize)
    # Headline and roofline: every (frame, sphere) pair of every obstacle leaf goes through the full arithmetic.  The
    # library defaults -- the exact early-out of the obstacle kernel, one pair loop for obstacle leaves that share their
    # control point -- are measured separately below.
    tree.set_early_out(False)
    tree.set_merge_coincident(False)

    n_buffers = 4 if O_ else 1
    q, qd, goal, spheres = S.synth_inputs_device(fk, n, B, O_, n_buffers, seed=S.SEEDS[config] + 17 * rank, device=device)
    goals = goal.reshape(B, 1, 3).contiguous()
    qdd = torch.empty(B, n, device=device)
    tree.reserve(B, O_)

    def step(i):
        tree.step(q, qd, qdd, goals=goals, spheres=spheres[i % n_buffers] if O_ else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    tree.profile(True)                      # CUDA events around every kernel launch of the timed region
    tree.profile_read()
    launches0 = _native.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        start.record()
        for i in range(args.steps):
            step(i)
        stop.record()
        barrier()
    elapsed_ms = max_over_ranks(start.elapsed_time(stop))
    launches = _native.launch_count() - launches0
    kernel_ms = tree.profile_read()         # {kernel: (total ms, launches)} on this rank
    tree.profile(False)
    ms_per_step = elapsed_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    per_gpu = B / (ms_per_step * 1e-3)

    # ---- same workload with the library default (exact early-out of pairs beyond the metric radius)
    def timed_mode(early_out, merged):
        """ms per step (max over ranks) and per-kernel CUDA-event times of the same workload in another mode."""
        tree.set_early_out(early_out)
        tree.set_merge_coincident(merged)
        for i in range(3):
            step(i)
        barrier()
        tree.profile(True)
        tree.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
        kms = tree.profile_read()
        tree.profile(False)
        return max_over_ranks(e0.elapsed_time(e1)), {k: v[0] / max(v[1], 1) for k, v in kms.items() if v[1]}

    early = library_default = None
    if O_ and not args.skip_early_out:
        ems, _ = timed_mode(True, False)
        sub = slice(0, min(B, 8192))
        frames = S.collision_frames(fk)
        origins = torch.stack([fk.forward(q[sub], fr)[:, :3, 3] for fr in frames], dim=1)          # [b,K,3]
        sp = spheres[0][sub]
        dist_s = torch.linalg.norm(origins[:, :, None, :] - sp[:, None, :, :3], dim=-1) - sp[:, None, :, 3]
        early = {"value": world * B / (ems / args.steps * 1e-3), "unit": UNIT, "ms_per_step": ems / args.steps,
                 "speedup_over_all_pairs": ms_per_step / (ems / args.steps),
                 "active_pair_fraction": float((dist_s.abs() <= 0.5).float().mean()),
                 "note": "library default: pairs beyond metric_modulation_radius contribute exactly zero "
                         "(reference rmp2.py:194) and are skipped; results are identical"}
        # ---- the library default: early-out and one pair loop per group of obstacle leaves with a common control point
        dms, dk = timed_mode(True, True)
        mms, mk = timed_mode(False, True)
        n_leaves, n_loops = tree.obstacle_slots()
        library_default = {
            "value": world * B / (dms / args.steps * 1e-3), "unit": UNIT, "ms_per_step": dms / args.steps,
            "speedup_over_all_pairs": ms_per_step / (dms / args.steps), "kernel_ms": dk,
            "obstacle_leaves": n_leaves, "pair_loops": n_loops,
            "all_pairs_merged": {"value": world * B / (mms / args.steps * 1e-3), "ms_per_step": mms / args.steps,
                                 "kernel_ms": mk},
            "note": "what rmp2_step does when no option is set: RMP2_OPT_EARLY_OUT and RMP2_OPT_MERGE_COINCIDENT on.  "
                    "Obstacle leaves with equal parameters on frames whose origins coincide for every q (Panda: joint2 "
                    "on joint1, joint6 on joint5) have identical pulled-back (M, f) -- the distance map differentiates "
                    "through the frame origin only (taskmap.py:124-128) -- so one of them runs the pair loop and its "
                    "sums are doubled; leaves whose control point cannot move (joint1 / joint2: J = 0, exactly zero "
                    "contribution in the reference as well) run none; `value` and the roofline above run every leaf's loop"}
        tree.set_early_out(False)
        tree.set_merge_coincident(False)

    # ---- result collection (north star: "NCCL allgather only for result collection"): not part of the step
    collect = None
    if world > 1:
        from riemannian_motion_policies_b200 import sharding
        gathered = torch.empty(world * B, n, device=device)
        for _ in range(2):
            sharding.gather_environments(qdd, world * B, out=gathered)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        g0.record()
        for _ in range(reps):
            sharding.gather_environments(qdd, world * B, out=gathered)
        g1.record()
        barrier()
        gms = max_over_ranks(g0.elapsed_time(g1)) / reps
        peer = (rank + 1) % world
        mine = bool(torch.equal(gathered[rank * B:(rank + 1) * B], qdd))
        # the neighbour's block must be what the neighbour computed: compare checksums
        sums = torch.stack([gathered[r * B:(r + 1) * B].double().sum() for r in range(world)])
        own = qdd.double().sum().reshape(1)
        all_own = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(all_own, own)
        ok = mine and bool(torch.equal(sums, torch.cat(all_own)))
        bytes_rank = B * n * 4
        collect = {"allgather_ms": gms, "bytes_per_rank": bytes_rank, "bytes_received_per_rank": (world - 1) * bytes_rank,
                   "gbs_per_rank": (world - 1) * bytes_rank / (gms * 1e-3) / 1e9,
                   "frac_of_nvlink5_900gbs": (world - 1) * bytes_rank / (gms * 1e-3) / 1e9 / 900.0,
                   "share_of_step": gms / ms_per_step, "blocks_match_owners": ok, "checked_peer": peer,
                   "note": "dist.all_gather_into_tensor (NCCL) of qdd [B,7] f32 into a preallocated [N*B,7]; outside the step"}

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.skip_e2e:
        hb = 2 if O_ else 1
        q_h, qd_h, goals_h = (t.cpu().pin_memory() for t in (q, qd, goals))
        sph_h = [spheres[i].cpu().pin_memory() for i in range(hb)] if O_ else [None]
        qdd_h = torch.empty(B, n).pin_memory()
        e2e_steps = max(20, min(args.steps, 50))
        for i in range(2):
            tree.step_host(q_h, qd_h, qdd_h, goals=goals_h, spheres=sph_h[i % hb])   # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            tree.step_host(q_h, qd_h, qdd_h, goals=goals_h, spheres=sph_h[i % hb])
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        h2d = B * (2 * n + 3 + 4 * O_) * 4
        d2h = B * n * 4
        peak = h2d_peak_gbs(torch, device, barrier)
        peak_min = -max_over_ranks(-peak)
        h2d_rate = h2d * e2e_steps / dt / 1e9
        e2e = {"value": world * B * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "h2d_gbs_per_rank": h2d_rate, "h2d_peak_gbs_per_rank": peak_min,
               "frac_of_h2d_peak": h2d_rate / peak_min,
               "h2d_peak_how": f"pinned 1 GiB cudaMemcpyAsync x5 per rank, {world} rank(s) copying concurrently, slowest rank",
               "note": "RmpCore -> CompiledTree.step_host -> rmp2_step_host, pinned host tensors, 64k-env chunks "
                       "pipelined over 3 streams; the step is bound by the host-to-device copy of the sphere rows",
               "host_cores_bound_per_rank": numa_cores}
        if not O_:
            e2e["matches_device_path"] = bool(torch.allclose(qdd_h[:4096], qdd[:4096].cpu(), rtol=0, atol=0))
        del q_h, qd_h, goals_h, sph_h, qdd_h

    parity = others = rollout = latency = cpu_baseline = cpu_vec = None
    if rank == 0 and world == 1 and not args.skip_checks:
        del spheres, q, qd, goal, goals, qdd
        torch.cuda.empty_cache()
        parity = fixture_parity(ns, config, n)
        if not args.skip_extras:
            others = other_config_lines(ns, S, fk, n, device, min(args.steps, 30), torch)
            rollout = rollout_lines(ns, S, fk, n, device, torch)
            latency = b1_latency(ns, S, fk, n, torch)
        cores = min(host_cores(), 64)
        per_core = max(args.envs_per_core, 24)
        v, total, wall = cpu_reference_throughput(config, n, cores, per_core, O_)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{total} envs of the same workload in rounds of {cores} processes x {per_core} envs "
                                  f"(median round, {wall:.1f} s measured after warm-up), oracle port run one env per "
                                  f"call like the reference"}
        cpu_vec = cpu_vectorised_baseline(config, n, cores)
        if latency is not None:
            latency["cpu_port_us_per_env_per_core"] = 1e6 * cores / v

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        info = tree.kernel_info(O_ if O_ else 64)
        t_s = ms_per_step * 1e-3
        gbs = BYTES_PER_ENV[config] * B / t_s / 1e9
        tflops = FLOPS_PER_ENV[config] * B / t_s / 1e12
        # dominant kernel: the obstacle pair loop when the tree has obstacles, else the step kernel
        dom = "spheres" if O_ else "step"
        n_slots = len(S.collision_frames(fk)) if O_ else 0
        dom_ms = kernel_ms[dom][0] / max(1, kernel_ms[dom][1])
        total_kernel_ms = sum(v[0] for v in kernel_ms.values()) / args.steps
        # algorithmic bytes / flops of one launch of the dominant kernel
        if O_:
            dom_bytes = B * (16.0 * O_ + 76.0 * n_slots)           # spheres in, frame record in (40 B; 36 B since round 2), 36 B (S,g) out
            dom_flops = B * 76.0 * O_ * n_slots                    # SURVEY.md 8d term C
        else:
            dom_bytes = B * BYTES_PER_ENV[config]
            dom_flops = B * FLOPS_PER_ENV[config]
        # measured DRAM traffic of the same kernel at the same size (one ncu --set full capture, profiles/)
        traffic, traffic_src = None, None
        for tname in ("r2_traffic.json", "r1_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if traffic is None and os.path.exists(tpath):
                with open(tpath) as fh:
                    tj = json.load(fh)
                if tj.get("config") == config and tj.get("envs") == B:
                    for kname, kv in tj["kernels"].items():
                        if kname.startswith(f"rmp2_{dom}_kernel"):
                            traffic, traffic_src = kv["dram_bytes_read"] + kv["dram_bytes_write"], "profiles/" + tname
        dom_gbs = dom_bytes / (dom_ms * 1e-3) / 1e9
        lanes = None
        if O_:
            pairs = float(B) * O_ * n_slots
            lane_tops = pairs * LANE_OPS_PER_PAIR / (dom_ms * 1e-3) / 1e12
            mufu_tops = pairs * MUFU_PER_PAIR / (dom_ms * 1e-3) / 1e12
            pipe = None
            ppath = os.path.join(ROOT, "profiles", "r1_pipe_peaks.json")
            if os.path.exists(ppath):
                with open(ppath) as fh:
                    pj = json.load(fh)
                pipe = {"ffma2_fma_per_clk_per_sm": pj.get("ffma2_fma_per_clk_per_sm"),
                        "ffma_per_clk_per_sm": pj.get("ffma_per_clk_per_sm"),
                        "mufu_per_clk_per_sm": (pj.get("mufu_per_clk_per_sm") or {}).get("ex2")}
            lanes = {"bound": "fp32+alu lane time", "lane_ops_per_pair": LANE_OPS_PER_PAIR, "mufu_per_pair": MUFU_PER_PAIR,
                     "achieved_tera_lane_ops": lane_tops, "peak": LANE_PEAK_TOPS, "frac": lane_tops / LANE_PEAK_TOPS,
                     "mufu_frac": mufu_tops / MUFU_PEAK_TOPS, "measured_pipe_rates": pipe,
                     "note": "executed (not algorithmic) thread-level FP32+ALU operations of the pair loop over "
                             "148 x 128 lanes x f_SM; ALU-pipe instructions cost FP32 lane time on B200 "
                             "(tools/pipe_peaks.cu), so this, not the flop count, is what the kernel saturates"}
        dom_tflops = dom_flops / (dom_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD[config], "envs_per_gpu": B, "global_envs": world * B, "spheres_per_env": O_,
                       "dof": n, "parallelism": f"env-sharded x{world}, no step-path collective",
                       "kernels": "frames/step rebuilt for this tree by NVRTC (rmp2_tree_specialize)" if specialized["on"]
                                  else "generic table-driven frames/step kernels",
                       "l2_policy": f"inputs larger than L2: {n_buffers} rotating sphere buffers of "
                                    f"{B * O_ * 16 / 1e6:.0f} MB each" if O_ else "q/qd/goal re-read each step",
                       "pairs": "every (obstacle leaf, sphere) pair through the full arithmetic: RMP2_OPT_EARLY_OUT and "
                                "RMP2_OPT_MERGE_COINCIDENT off for value / roofline / e2e; the library defaults are "
                                "timed under early_out and library_default" if O_ else "no obstacle leaves"},
            "roofline": {"bound": "hbm", "kernel": f"rmp2_{dom}_kernel", "achieved": dom_gbs, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": dom_gbs / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_kind,
                         "kernel_ms_per_launch": dom_ms, "kernel_share_of_step": dom_ms / max(total_kernel_ms, 1e-12),
                         "algorithmic_bytes_per_launch": dom_bytes,
                         "note": "this path is FP32/MUFU-issue bound by design, not HBM bound (see roofline_fp32); the "
                                 "HBM figure is given because the contract's bounds are hbm|tensor"},
            "roofline_fp32": {"bound": "fp32", "kernel": f"rmp2_{dom}_kernel", "achieved": dom_tflops,
                              "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": dom_tflops / FP32_PEAK_TFLOPS,
                              "algorithmic_flops_per_launch": dom_flops,
                              "whole_step": {"achieved": tflops, "frac": tflops / FP32_PEAK_TFLOPS,
                                             "flops_per_env_step": FLOPS_PER_ENV[config],
                                             "hbm_gbs": gbs, "bytes_per_env_step": BYTES_PER_ENV[config]},
                              "peak_source": "analytic 148 SM x 128 lanes x 2 x 1.965 GHz",
                              "lanes": lanes},
            "kernel_ms": {k: {"ms_per_step": v[0] / args.steps, "launches": int(v[1])} for k, v in kernel_ms.items()},
            "per_gpu_value": per_gpu, "gpu_launches": int(launches), "kernel": info, "specialized": specialized,
            "clocks": clocks.summary(),
            "early_out": early, "library_default": library_default, "e2e": e2e, "collect": collect, "cpu_baseline": cpu_baseline,
            "cpu_baseline_vectorised": cpu_vec, "parity": parity, "other_configs": others, "rollout": rollout,
            "latency_b1_us": latency,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4, 5])
    ap.add_argument("--envs", type=int, default=0, help="environments per GPU (default: 1,048,576)")
    ap.add_argument("--envs-per-core", type=int, default=6, help="CPU arm: environments per host process per step")
    ap.add_argument("--reference-budget-s", type=float, default=150.0, help="CPU arm: stop after this many seconds")
    ap.add_argument("--no-specialize", action="store_true", help="keep the generic (table-interpreting) frames/step kernels")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-early-out", action="store_true")
    ap.add_argument("--skip-checks", action="store_true", help="skip the parity block, the CPU baselines and the extras")
    ap.add_argument("--skip-extras", action="store_true", help="skip other_configs / rollout / latency_b1_us (N = 1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
