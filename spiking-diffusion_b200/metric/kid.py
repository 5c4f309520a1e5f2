"""Kernel inception distance on given features.

R/main.py:465-490 calls ``torchmetrics.image.kid.KernelInceptionDistance()`` (defaults: subsets=100, subset_size=1000,
degree=3, gamma=None -> 1/d, coef=1).  torchmetrics is a third-party dependency that is not part of the reference tree
(and no version is pinned there); its published estimator is restated: ``poly_mmd`` = unbiased MMD^2 under the polynomial
kernel, averaged over random subsets.  The MMD is one CUDA kernel family (csrc/metrics.cu); the subset draw is host logic.
"""
import torch

from .._lib import check, lib, on_device_of, ptr, stream_ptr
from .common import require_cuda, workspace


@on_device_of
def poly_mmd(f_real: torch.Tensor, f_fake: torch.Tensor, degree: int = 3, gamma=None, coef: float = 1.0) -> torch.Tensor:
    """Unbiased MMD^2 of two feature sets of equal size [m, d] under k(x, y) = (gamma x.y + coef)^degree."""
    require_cuda(f_real, f_fake)
    if f_real.dim() != 2 or f_real.shape != f_fake.shape or f_real.shape[0] < 2:
        raise ValueError("poly_mmd expects two [m >= 2, d] feature matrices of one shape")
    x, y = f_real.float().contiguous(), f_fake.float().contiguous()
    m, d = x.shape
    g = 1.0 / d if gamma is None else float(gamma)
    out = torch.empty((), dtype=torch.float64, device=x.device)
    ws = workspace(x.device, m=m)
    check(lib().sd_metric_poly_mmd2(ptr(x), ptr(y), m, d, int(degree), g, float(coef), ptr(out), ptr(ws), stream_ptr()))
    return out


def kernel_inception_distance_from_features(f_real, f_fake, subsets: int = 100, subset_size: int = 1000, degree: int = 3,
                                            gamma=None, coef: float = 1.0, generator=None):
    """(mean, std) of poly_mmd over ``subsets`` random subsets of ``subset_size`` rows of each set, drawn with
    torch.randperm as torchmetrics' KernelInceptionDistance.compute does."""
    n_r, n_f = f_real.shape[0], f_fake.shape[0]
    if subset_size > n_r or subset_size > n_f:
        raise ValueError("Argument `subset_size` should be smaller than the number of samples")
    scores = []
    for _ in range(subsets):
        pr = torch.randperm(n_r, generator=generator)[:subset_size].to(f_real.device)
        pf = torch.randperm(n_f, generator=generator)[:subset_size].to(f_fake.device)
        scores.append(poly_mmd(f_real[pr], f_fake[pf], degree, gamma, coef))
    s = torch.stack(scores)
    return s.mean(), s.std(unbiased=False)
