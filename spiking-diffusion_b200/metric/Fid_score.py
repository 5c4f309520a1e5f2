"""Frechet distance on Inception features (the network-free half of R/metric/Fid_score.py).

``calculate_activation_statistics_from_features(act)`` is R/metric/Fid_score.py:100-113 from the point where the
activations exist; ``calculate_frechet_distance(mu1, sigma1, mu2, sigma2)`` is :116-173 including the reference's own
``sqrtm`` (:14-17: U diag(sqrt(S)) Vh of an SVD, not scipy's matrix square root).  fp64 on the GPU."""
import ctypes

import torch

from .._lib import check, lib, on_device_of, ptr, stream_ptr
from .common import require_cuda, workspace


@on_device_of
def calculate_activation_statistics_from_features(act: torch.Tensor):
    """act: [N, d] fp32 or fp64 CUDA tensor -> (mu [d], sigma [d, d]) fp64: np.mean(act, 0), np.cov(act, rowvar=False)."""
    require_cuda(act)
    if act.dim() != 2 or act.shape[0] < 2:
        raise ValueError("expected activations of shape [N >= 2, d]")
    if act.dtype not in (torch.float32, torch.float64):
        act = act.float()
    act = act.contiguous()
    N, d = act.shape
    mu = torch.empty(d, dtype=torch.float64, device=act.device)
    sigma = torch.empty((d, d), dtype=torch.float64, device=act.device)
    check(lib().sd_metric_feature_stats(ptr(act), int(act.dtype == torch.float64), N, d, ptr(mu), ptr(sigma), stream_ptr()))
    return mu, sigma


@on_device_of
def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6, return_sweeps=False):
    """d^2 = ||mu_1 - mu_2||^2 + Tr(C_1 + C_2 - 2 sqrtm(C_1 C_2)) with the reference's SVD-based sqrtm.
    ``eps`` is accepted for signature compatibility: the reference only uses it when its sqrtm returns non-finite values,
    which an SVD of finite inputs does not."""
    require_cuda(mu1, sigma1, mu2, sigma2)
    mu1, mu2 = mu1.reshape(-1).double().contiguous(), mu2.reshape(-1).double().contiguous()
    sigma1, sigma2 = sigma1.double().contiguous(), sigma2.double().contiguous()
    if sigma1.dim() == 0:                       # np.atleast_2d
        sigma1, sigma2 = sigma1.reshape(1, 1), sigma2.reshape(1, 1)
    assert mu1.shape == mu2.shape, "Training and test mean vectors have different lengths"
    assert sigma1.shape == sigma2.shape, "Training and test covariances have different dimensions"
    d = mu1.numel()
    if sigma1.shape != (d, d):
        raise ValueError(f"covariances must be [{d}, {d}]")
    out = torch.empty((), dtype=torch.float64, device=mu1.device)
    ws = workspace(mu1.device, d=d)
    sweeps = ctypes.c_int(0)
    check(lib().sd_metric_frechet(ptr(mu1), ptr(sigma1), ptr(mu2), ptr(sigma2), d, ptr(out), ctypes.byref(sweeps), ptr(ws),
                                  stream_ptr()))
    return (out, sweeps.value) if return_sweeps else out


def calculate_fid_from_features(act1: torch.Tensor, act2: torch.Tensor) -> torch.Tensor:
    """calculate_fid (R/metric/Fid_score.py:226-245) from the point where both activation matrices exist."""
    mu1, s1 = calculate_activation_statistics_from_features(act1)
    mu2, s2 = calculate_activation_statistics_from_features(act2)
    return calculate_frechet_distance(mu1, s1, mu2, s2)
