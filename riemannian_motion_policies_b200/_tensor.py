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


def like_input(result, reference_input):
    """Return ``result`` on the side (host/device) the caller's input lived on."""
    if is_device_tensor(unwrap(reference_input)):
        return result
    return result.cpu()


def current_stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream
