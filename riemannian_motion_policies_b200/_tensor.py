"""Tensor plumbing shared by the host-side mirror classes.

Rule of the package: *host in -> host out, device in -> device out*.  Inputs given the way the
reference's callers give them (NumPy arrays, lists, CPU tensors) are staged to the GPU, the CUDA
kernels run, and the result comes back as a CPU tensor, so ``core.evaluate(q, qd).numpy()`` keeps
working (reference: experiments/franka_panda/05_obstacle_avoidance.py:96).  CUDA tensors stay on
the GPU.  There is no CPU compute path: without a CUDA device every call below raises.
"""
import numpy as np
import torch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("riemannian_motion_policies_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def is_device_tensor(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def unwrap(x):
    """Accept Variable-like holders (``.value()``) the way TF variables are accepted."""
    if hasattr(x, "value") and callable(x.value) and not isinstance(x, torch.Tensor):
        return x.value()
    return x


def to_device(x, device=None, dtype=torch.float32):
    """Any array-like -> contiguous float32 CUDA tensor."""
    x = unwrap(x)
    device = device or require_cuda()
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=dtype).to(device).contiguous()


def to_host_array(x):
    """Any host array-like -> float32 ndarray (the value a float32 CUDA tensor made by ``to_device`` would hold)."""
    x = unwrap(x)
    if isinstance(x, torch.Tensor):
        return x.detach().to(dtype=torch.float32).numpy()
    return np.asarray(x, dtype=np.float64).astype(np.float32)


def stage_host_inputs(arrays, device):
    """Host array-likes (or None) -> float32 CUDA tensors of the same shapes with ONE host-to-device copy: the
    reference's call pattern is one environment per ``evaluate`` (05_obstacle_avoidance.py:96), where a copy per
    argument costs more than the kernels.  Every section starts on a 16-byte boundary (sphere rows need it)."""
    host = [None if a is None else to_host_array(a) for a in arrays]
    offsets, total = [], 0
    for h in host:
        offsets.append(total)
        if h is not None:
            total += (h.size + 3) & ~3
    flat = np.zeros(max(total, 4), dtype=np.float32)
    for h, off in zip(host, offsets):
        if h is not None:
            flat[off:off + h.size] = h.reshape(-1)
    dev = torch.from_numpy(flat).to(device)
    return [None if h is None else dev[off:off + h.size].view(h.shape) for h, off in zip(host, offsets)]


def like_input(result, reference_input):
    """Return ``result`` on the side (host/device) the caller's input lived on."""
    if is_device_tensor(unwrap(reference_input)):
        return result
    return result.cpu()


def current_stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream
