"""Timing of the quality-metric kernels at the reference's sizes (GPU box): Frechet distance on 2048-d features
(1 280 real / 1 000 generated samples as in R/main.py:492-529), KID subsets, SSIM / MSE on a 10 000-image test set."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import metrics_oracle as M  # noqa: E402
from spiking_diffusion_b200 import metric  # noqa: E402


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts), r


def main():
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    f1 = torch.from_numpy(M.synth_features(1, 1280, d).astype(np.float32)).cuda()
    f2 = torch.from_numpy(M.synth_features(2, 1000, d, 0.3).astype(np.float32)).cuda()
    t, (mu1, s1) = timed(lambda: metric.Fid_score.calculate_activation_statistics_from_features(f1))
    print(f"feature statistics N=1280 d={d}: {t * 1e3:.2f} ms")
    mu2, s2 = metric.Fid_score.calculate_activation_statistics_from_features(f2)
    t, (fid, sweeps) = timed(lambda: metric.Fid_score.calculate_frechet_distance(mu1, s1, mu2, s2, return_sweeps=True), reps=1)
    t0 = time.perf_counter()
    ref = M.frechet_distance(mu1.cpu().numpy(), s1.cpu().numpy(), mu2.cpu().numpy(), s2.cpu().numpy())
    t_cpu = time.perf_counter() - t0
    print(f"Frechet distance d={d}: {t:.3f} s on the GPU ({sweeps} Jacobi sweeps), value {float(fid):.9f}; "
          f"numpy (the reference's SVD) {t_cpu:.2f} s, value {ref:.9f}, rel. diff {abs(float(fid) - ref) / abs(ref):.2e}")
    x, y = f1[:1000].contiguous(), f2[:1000].contiguous()
    t, v = timed(lambda: metric.kid.poly_mmd(x, y))
    print(f"poly_mmd m=1000 d={d}: {t * 1e3:.2f} ms (value {float(v):.6e}; oracle {M.poly_mmd(x.cpu().numpy(), y.cpu().numpy()):.6e})")
    a, b = M.synth_images(3, 10000, 1, 28, 28)
    a, b = a.cuda(), b.cuda()
    t, v = timed(lambda: metric.pytorch_ssim.SSIM(window_size=11)(a, b))
    print(f"SSIM 10000 x 1 x 28 x 28: {t * 1e3:.2f} ms; MSE: {timed(lambda: metric.mse_loss(a, b))[0] * 1e3:.2f} ms")


if __name__ == "__main__":
    main()
