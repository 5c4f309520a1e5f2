"""Golden vectors for the quality-metric kernels -- run in the BUILD container (needs /root/reference or baseline/_ref).

For every case of oracle/metrics_oracle.py:METRIC_CASES the seeded inputs go through the UNMODIFIED reference functions
(metric.pytorch_ssim.SSIM / ssim, metric.Fid_score.calculate_frechet_distance) and through the oracle restatement, which
must agree (bit-identical for SSIM, 1e-12 relative for the Frechet distance); the MMD oracle is cross-checked against
scikit-learn's polynomial_kernel; the inception score calls scipy.stats.entropy itself.  Only the OUTPUTS are stored
(tests regenerate the inputs from the seeds): tests/golden/metrics.npz.

    python oracle/gen_golden_metrics.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import metrics_oracle as M, ref_loader  # noqa: E402


def main():
    ref = ref_loader.load_metrics()
    out = {}
    for i, c in enumerate(M.METRIC_CASES["ssim"]):
        a, b = M.synth_images(c["seed"], c["N"], c["C"], c["H"], c["W"])
        r_mean = ref.pytorch_ssim.SSIM(window_size=c["ws"])(a, b)
        r_per = ref.pytorch_ssim.ssim(a, b, window_size=c["ws"], size_average=False)
        o_mean, o_per = M.ssim(a, b, c["ws"]), M.ssim(a, b, c["ws"], size_average=False)
        assert torch.equal(r_mean, o_mean) and torch.equal(r_per, o_per), "SSIM oracle differs from the reference"
        out[f"ssim{i}_mean"], out[f"ssim{i}_per"] = r_mean.numpy(), r_per.numpy()
        out[f"mse{i}"] = np.float32(M.mse(a, b))
    for i, c in enumerate(M.METRIC_CASES["frechet"]):
        f1 = M.synth_features(c["seed"], c["N1"], c["d"], 0.0, c.get("rank"))
        f2 = M.synth_features(c["seed"] + 100, c["N2"], c["d"], c["shift"], c.get("rank"))
        mu1, s1 = M.feature_stats(f1)
        mu2, s2 = M.feature_stats(f2)
        r = float(ref.calculate_frechet_distance(mu1, s1, mu2, s2))
        o = M.frechet_distance(mu1, s1, mu2, s2)
        assert abs(r - o) <= 1e-12 * max(1.0, abs(r)), (r, o)
        out[f"fid{i}"], out[f"fid{i}_mu1"], out[f"fid{i}_tr1"] = np.float64(r), mu1, np.float64(np.trace(s1))
    from sklearn.metrics.pairwise import polynomial_kernel
    for i, c in enumerate(M.METRIC_CASES["mmd"]):
        x = M.synth_features(c["seed"], c["m"], c["d"]).astype(np.float32)
        y = M.synth_features(c["seed"] + 100, c["m"], c["d"], c["shift"]).astype(np.float32)
        o = M.poly_mmd(x, y)
        m = c["m"]
        kxx, kyy, kxy = (polynomial_kernel(p.astype(np.float64), q.astype(np.float64), degree=3, gamma=1.0 / c["d"], coef0=1.0)
                         for p, q in ((x, x), (y, y), (x, y)))
        chk = ((kxx.sum() - np.trace(kxx)) + (kyy.sum() - np.trace(kyy))) / (m * (m - 1)) - 2 * kxy.sum() / m ** 2
        assert abs(o - chk) <= 1e-10 * max(1.0, abs(o)), (o, chk)
        out[f"mmd{i}"] = np.float64(o)
    for i, c in enumerate(M.METRIC_CASES["is"]):
        mean, std = M.inception_score(M.synth_probs(c["seed"], c["N"], c["K"]), c["splits"])
        out[f"is{i}"] = np.array([mean, std])
    path = os.path.join(ROOT, "tests", "golden", "metrics.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path)} bytes, {len(out)} arrays) from the reference at {ref_loader.source()}")


if __name__ == "__main__":
    main()
