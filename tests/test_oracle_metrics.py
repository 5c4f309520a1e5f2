"""Quality-metric oracle (oracle/metrics_oracle.py) against the committed outputs of the unmodified reference
(tests/golden/metrics.npz, oracle/gen_golden_metrics.py) and, where the reference is mounted or installed, against the
reference functions themselves.  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import metrics_oracle as M, ref_loader


def test_ssim_and_mse_oracle_match_reference_outputs():
    g = golden("metrics.npz")
    for i, c in enumerate(M.METRIC_CASES["ssim"]):
        a, b = M.synth_images(c["seed"], c["N"], c["C"], c["H"], c["W"])
        assert np.array_equal(M.ssim(a, b, c["ws"]).numpy(), g[f"ssim{i}_mean"])
        assert np.array_equal(M.ssim(a, b, c["ws"], size_average=False).numpy(), g[f"ssim{i}_per"])
        assert np.float32(M.mse(a, b)) == g[f"mse{i}"]
    a, _ = M.synth_images(0, 2, 1, 28, 28)
    assert abs(float(M.ssim(a, a)) - 1.0) < 1e-6            # known answer: identical images


def test_frechet_oracle_matches_reference_outputs():
    g = golden("metrics.npz")
    for i, c in enumerate(M.METRIC_CASES["frechet"]):
        f1 = M.synth_features(c["seed"], c["N1"], c["d"], 0.0, c.get("rank"))
        f2 = M.synth_features(c["seed"] + 100, c["N2"], c["d"], c["shift"], c.get("rank"))
        mu1, s1 = M.feature_stats(f1)
        mu2, s2 = M.feature_stats(f2)
        assert np.allclose(mu1, g[f"fid{i}_mu1"], rtol=0, atol=1e-14) and abs(np.trace(s1) - g[f"fid{i}_tr1"]) < 1e-10
        got = M.frechet_distance(mu1, s1, mu2, s2)
        assert abs(got - float(g[f"fid{i}"])) <= 1e-9 * max(1.0, abs(got))
    mu, s = M.feature_stats(M.synth_features(1, 100, 16))
    assert abs(M.frechet_distance(mu, s, mu, s)) < 1e-9     # known answer: identical statistics


def test_mmd_and_inception_score_oracle_match_golden():
    g = golden("metrics.npz")
    for i, c in enumerate(M.METRIC_CASES["mmd"]):
        x = M.synth_features(c["seed"], c["m"], c["d"]).astype(np.float32)
        y = M.synth_features(c["seed"] + 100, c["m"], c["d"], c["shift"]).astype(np.float32)
        assert abs(M.poly_mmd(x, y) - float(g[f"mmd{i}"])) <= 1e-10 * max(1.0, abs(float(g[f"mmd{i}"])))
    for i, c in enumerate(M.METRIC_CASES["is"]):
        mean, std = M.inception_score(M.synth_probs(c["seed"], c["N"], c["K"]), c["splits"])
        assert np.allclose([mean, std], g[f"is{i}"], rtol=1e-12, atol=1e-14)
    # known answers: one-hot rows over K classes, uniformly spread -> IS = K; identical rows -> IS = 1
    K = 10
    assert abs(M.inception_score(np.eye(K)[np.arange(100) % K] * 1.0, 1)[0] - K) < 1e-9
    assert abs(M.inception_score(np.tile(M.synth_probs(0, 1, K), (20, 1)), 2)[0] - 1.0) < 1e-12


@pytest.mark.skipif(not ref_loader.available(), reason="reference neither mounted nor installed")
def test_oracle_pinned_to_the_live_reference_functions():
    ref = ref_loader.load_metrics()
    a, b = M.synth_images(21, 4, 3, 28, 28)
    assert torch.equal(ref.pytorch_ssim.SSIM(window_size=11)(a, b), M.ssim(a, b, 11))
    assert torch.equal(ref.pytorch_ssim.ssim(a, b, 5, size_average=False), M.ssim(a, b, 5, size_average=False))
    mu1, s1 = M.feature_stats(M.synth_features(22, 90, 40))
    mu2, s2 = M.feature_stats(M.synth_features(23, 70, 40, 0.2))
    r = float(ref.calculate_frechet_distance(mu1, s1, mu2, s2))
    assert abs(r - M.frechet_distance(mu1, s1, mu2, s2)) <= 1e-12 * max(1.0, abs(r))
    A = s1.dot(s2)
    assert np.allclose(ref.sqrtm(A), M.sqrtm_svd(A), rtol=0, atol=1e-12)
