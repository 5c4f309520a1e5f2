"""CPU oracle of the quality-metric algebra -- TEST INFRASTRUCTURE ONLY.

numpy / torch-CPU restatements of the reference's evaluation arithmetic, each citing the lines it follows.  Pin status:
``ssim`` and ``frechet_distance`` are PINNED bit-for-bit / to 1e-12 against the unmodified reference functions
(``tests/test_oracle_metrics.py`` through ``oracle/ref_loader.load_metrics()`` whenever the reference is mounted or
installed, and through the committed ``tests/golden/metrics.npz`` otherwise); ``inception_score`` calls the very
``scipy.stats.entropy`` the reference calls; ``poly_mmd`` restates torchmetrics' published estimator (torchmetrics is a
third-party dependency absent from the reference tree and from this image, unpinned in the reference) and is
cross-checked against scikit-learn's ``polynomial_kernel``."""
from math import exp

import numpy as np
import torch
import torch.nn.functional as F


def mse(a: torch.Tensor, b: torch.Tensor) -> float:
    """R/main.py:319: F.mse_loss(recon_images, norm_images).item()."""
    return float(F.mse_loss(a, b))


def _window(window_size: int, channel: int) -> torch.Tensor:
    # R/metric/pytorch_ssim/__init__.py:7-15
    g = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True):
    """R/metric/pytorch_ssim/__init__.py:17-37 (_ssim) with the window of :7-15."""
    ch = img1.shape[1]
    w = _window(window_size, ch)
    pad = window_size // 2
    mu1 = F.conv2d(img1, w, padding=pad, groups=ch)
    mu2 = F.conv2d(img2, w, padding=pad, groups=ch)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=pad, groups=ch) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=pad, groups=ch) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=pad, groups=ch) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean() if size_average else m.mean(1).mean(1).mean(1)


def feature_stats(act: np.ndarray):
    """R/metric/Fid_score.py:110-112: mu = np.mean(act, axis=0); sigma = np.cov(act, rowvar=False)."""
    return np.mean(act, axis=0), np.cov(act, rowvar=False)


def sqrtm_svd(A: np.ndarray) -> np.ndarray:
    """R/metric/Fid_score.py:14-17: the reference's own 'sqrtm' -- U diag(sqrt(S)) Vh of the SVD."""
    U, S, V = np.linalg.svd(A)
    return U.dot(np.diag(np.sqrt(S))).dot(V)


def frechet_distance(mu1, sigma1, mu2, sigma2) -> float:
    """R/metric/Fid_score.py:116-173 (finite case)."""
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    diff = mu1 - mu2
    covmean = sqrtm_svd(sigma1.dot(sigma2))
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


def poly_mmd(f_real: np.ndarray, f_fake: np.ndarray, degree: int = 3, gamma=None, coef: float = 1.0) -> float:
    """torchmetrics.image.kid: poly_kernel (f1 @ f2.T * gamma + coef) ** degree with gamma = 1 / d by default;
    maximum_mean_discrepancy: (sum offdiag k_xx + sum offdiag k_yy) / (m (m-1)) - 2 sum k_xy / m^2."""
    x, y = f_real.astype(np.float64), f_fake.astype(np.float64)
    m, d = x.shape
    g = 1.0 / d if gamma is None else gamma
    kxx, kyy, kxy = (x @ x.T * g + coef) ** degree, (y @ y.T * g + coef) ** degree, (x @ y.T * g + coef) ** degree
    kt_xx = kxx.sum() - np.trace(kxx)
    kt_yy = kyy.sum() - np.trace(kyy)
    return float((kt_xx + kt_yy) / (m * (m - 1)) - 2 * kxy.sum() / (m ** 2))


def inception_score(preds: np.ndarray, splits: int = 1):
    """R/metric/IS_score.py:58-72 from the predictions on."""
    from scipy.stats import entropy
    N = preds.shape[0]
    split_scores = []
    for k in range(splits):
        part = preds[k * (N // splits): (k + 1) * (N // splits), :]
        py = np.mean(part, axis=0)
        scores = [entropy(part[i, :], py) for i in range(part.shape[0])]
        split_scores.append(np.exp(np.mean(scores)))
    return float(np.mean(split_scores)), float(np.std(split_scores))


# ---- seeded synthetic inputs shared by the generator script and the tests -----------------------------------------
def synth_images(seed: int, N: int, C: int, H: int, W: int):
    g = torch.Generator().manual_seed(seed)
    a = torch.rand((N, C, H, W), generator=g) - 0.5
    b = (a + 0.15 * torch.randn((N, C, H, W), generator=g)).clamp(-0.5, 0.5)
    return a, b


def synth_features(seed: int, N: int, d: int, shift: float = 0.0, rank=None):
    rng = np.random.default_rng(seed)
    r = d if rank is None else rank
    mix = rng.standard_normal((r, d)) / np.sqrt(r)
    return (rng.standard_normal((N, r)) @ mix + shift * rng.standard_normal(d)).astype(np.float64)


def synth_probs(seed: int, N: int, K: int):
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((N, K)) * 2.0
    e = np.exp(z - z.max(1, keepdims=True))
    return e / e.sum(1, keepdims=True)


METRIC_CASES = {
    "ssim": [dict(seed=1, N=6, C=1, H=28, W=28, ws=11), dict(seed=2, N=3, C=3, H=32, W=32, ws=11),
             dict(seed=3, N=2, C=1, H=9, W=13, ws=7)],
    "frechet": [dict(seed=4, N1=300, N2=280, d=64, shift=0.3), dict(seed=5, N1=40, N2=50, d=96, shift=0.1),     # N < d: singular
                dict(seed=6, N1=200, N2=200, d=33, shift=0.0), dict(seed=7, N1=500, N2=400, d=128, shift=0.5, rank=20)],
    "mmd": [dict(seed=8, m=100, d=64, shift=0.2), dict(seed=9, m=77, d=200, shift=0.0), dict(seed=10, m=256, d=48, shift=1.0)],
    "is": [dict(seed=11, N=200, K=50, splits=4), dict(seed=12, N=64, K=1000, splits=1)],
}
