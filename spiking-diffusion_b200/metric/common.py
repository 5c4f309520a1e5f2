"""Shared plumbing of the metric kernels: workspace, argument checks, F.mse_loss."""
import torch

from .._lib import check, lib, on_device_of, ptr, stream_ptr


def workspace(device, n_elements: int = 0, d: int = 0, m: int = 0) -> torch.Tensor:
    return torch.empty(lib().sd_metric_workspace_bytes(int(n_elements), int(d), int(m)), dtype=torch.uint8, device=device)


def require_cuda(*ts):
    for t in ts:
        if not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise RuntimeError("the metric kernels need CUDA tensors: spiking_diffusion_b200 has no CPU path")


@on_device_of
def mse_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """F.mse_loss(a, b) (mean reduction) as used by the reconstruction test, R/main.py:319.  0-d fp32 CUDA tensor."""
    require_cuda(a, b)
    if a.shape != b.shape:
        raise ValueError(f"mse_loss: shapes differ: {tuple(a.shape)} vs {tuple(b.shape)}")
    a, b = a.contiguous().float(), b.contiguous().float()
    out = torch.empty((), dtype=torch.float32, device=a.device)
    ws = workspace(a.device)
    check(lib().sd_metric_mse(ptr(a), ptr(b), a.numel(), ptr(out), ptr(ws), stream_ptr()))
    return out
