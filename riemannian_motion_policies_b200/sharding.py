"""Environment-sharded data parallelism: one process per GPU, no collective on the step path.

Environments are independent, so a batch of B environments is cut into `world_size` contiguous
blocks; every rank runs the same compiled tree on its block.  The only
communication is an optional all-gather of `qdd` when one host needs all results; it is not part
of the step and is timed separately.  Works with the `nccl` backend on GPUs and with `gloo` on CPU
tensors (used by the CPU tests of the partition / collection logic).
"""
import os

import torch
import torch.distributed as dist


def bind_host_to_gpu(device_index):
    """Pin this process (and the host memory it touches from now on: first-touch NUMA placement, which is
    what decides where pinned staging buffers live) to the CPU cores closest to GPU `device_index`.
    With one process per GPU on a two-socket box this keeps every rank's host<->device copies on its own
    socket's memory controllers and PCIe root instead of funnelling all eight through one socket.
    Returns the number of cores bound to, or 0 when NVML / sched_setaffinity is unavailable (no-op)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cores = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return 0
        os.sched_setaffinity(0, allowed)
        return len(allowed)
    except Exception:
        return 0


def shard_bounds(B, world_size, rank):
    """[lo, hi) of rank's contiguous block; the first B % world_size ranks hold one extra environment."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(int(B), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t, world_size, rank):
    """This rank's block of a [B, ...] tensor (a view, no copy)."""
    lo, hi = shard_bounds(t.shape[0], world_size, rank)
    return t[lo:hi]


def gather_environments(local, B, group=None, out=None):
    """All-gather per-rank blocks [b_r, ...] into the full [B, ...] tensor on every rank (NCCL on GPUs: result
    collection only, never on the step path).  Equal blocks (B divisible by the world size) gather straight into
    `out` (allocated when not given) with one collective and no staging copy; blocks that differ by one row are
    padded to the largest block for the collective."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(B, world, r) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    tail = tuple(local.shape[1:])
    if all(hi - lo == longest for lo, hi in sizes) and local.is_contiguous():
        if out is None:
            out = torch.empty((B,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    padded = torch.zeros((longest,) + tail, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((world * longest,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    buf = buf.reshape((world, longest) + tail)
    res = torch.cat([buf[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
    if out is not None:
        out.copy_(res)
        return out
    return res


class ShardedStep:
    """Runs `step_fn(q, qd, goals, spheres) -> qdd` on this rank's block of every input."""

    def __init__(self, step_fn, group=None):
        self.step_fn = step_fn
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def local_inputs(self, *tensors):
        return [None if t is None else shard(t, self.world, self.rank) for t in tensors]

    def step_local(self, q, qd, goals=None, spheres=None):
        """Inputs are the FULL batch (e.g. generated identically on every rank); only this rank's
        block is evaluated.  No communication."""
        ql, qdl, gl, sl = self.local_inputs(q, qd, goals, spheres)
        return self.step_fn(ql, qdl, gl, sl)

    def step_and_gather(self, q, qd, goals=None, spheres=None):
        local = self.step_local(q, qd, goals, spheres)
        if self.world == 1:
            return local
        return gather_environments(local, q.shape[0], self.group)
