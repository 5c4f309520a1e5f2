"""(f4) quality-metric kernels on the GPU against the CPU oracle on the same seeded inputs and against the committed
outputs of the unmodified reference (tests/golden/metrics.npz).  Tolerances: fp32 kernels (MSE, SSIM) 2e-6 absolute on
values of order 0.01..1 (the reference's own conv2d summation order is not specified); fp64 kernels 1e-9 relative."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import metrics_oracle as M
from spiking_diffusion_b200 import metric

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("i", range(len(M.METRIC_CASES["ssim"])))
def test_ssim_and_mse_match_oracle_and_reference_golden(i):
    c, g = M.METRIC_CASES["ssim"][i], golden("metrics.npz")
    a, b = M.synth_images(c["seed"], c["N"], c["C"], c["H"], c["W"])
    ac, bc = a.cuda(), b.cuda()
    got = metric.pytorch_ssim.SSIM(window_size=c["ws"])(ac, bc)
    assert got.shape == () and got.dtype == torch.float32
    assert abs(float(got) - float(M.ssim(a, b, c["ws"]))) <= 2e-6
    assert abs(float(got) - float(g[f"ssim{i}_mean"])) <= 2e-6
    per = metric.pytorch_ssim.ssim(ac, bc, window_size=c["ws"], size_average=False).cpu().numpy()
    assert per.shape == (c["N"],) and np.abs(per - g[f"ssim{i}_per"]).max() <= 2e-6
    mse = metric.mse_loss(ac, bc)
    assert abs(float(mse) - float(g[f"mse{i}"])) <= 1e-7 * max(1.0, float(g[f"mse{i}"])) + 1e-9
    # the recon test's loss: 1 - SSIM   (R/main.py:321)
    assert abs((1 - float(got)) - (1 - float(g[f"ssim{i}_mean"]))) <= 2e-6
    assert abs(float(metric.pytorch_ssim.ssim(ac, ac)) - 1.0) <= 1e-6


@pytest.mark.parametrize("i", range(len(M.METRIC_CASES["frechet"])))
def test_frechet_distance_matches_oracle_and_reference_golden(i):
    c, g = M.METRIC_CASES["frechet"][i], golden("metrics.npz")
    f1 = M.synth_features(c["seed"], c["N1"], c["d"], 0.0, c.get("rank"))
    f2 = M.synth_features(c["seed"] + 100, c["N2"], c["d"], c["shift"], c.get("rank"))
    mu1, s1 = metric.Fid_score.calculate_activation_statistics_from_features(torch.from_numpy(f1).cuda())
    mu2, s2 = metric.Fid_score.calculate_activation_statistics_from_features(torch.from_numpy(f2).cuda())
    o_mu1, o_s1 = M.feature_stats(f1)
    assert np.abs(mu1.cpu().numpy() - o_mu1).max() <= 1e-13 and np.abs(s1.cpu().numpy() - o_s1).max() <= 1e-12
    fid, sweeps = metric.Fid_score.calculate_frechet_distance(mu1, s1, mu2, s2, return_sweeps=True)
    ref = float(g[f"fid{i}"])
    assert 1 <= sweeps < 60, sweeps
    assert abs(float(fid) - ref) <= 1e-9 * max(1.0, abs(ref)), (float(fid), ref, sweeps)
    # fp32 activations (what the Inception network produces): statistics still accumulate in fp64
    mu1f, s1f = metric.Fid_score.calculate_activation_statistics_from_features(torch.from_numpy(f1).float().cuda())
    of = M.feature_stats(f1.astype(np.float32).astype(np.float64))
    assert np.abs(mu1f.cpu().numpy() - of[0]).max() <= 1e-12 and np.abs(s1f.cpu().numpy() - of[1]).max() <= 1e-11
    same = metric.Fid_score.calculate_frechet_distance(mu1, s1, mu1, s1)
    # identical statistics -> 0 up to sqrt(eps) * scale: the square roots of the (numerically) zero singular values of
    # a rank-deficient product are only accurate to that, in LAPACK as well as here
    assert abs(float(same)) <= 1e-7 * max(1.0, float(np.trace(o_s1)))


def test_frechet_from_features_end_to_end_and_argument_checks():
    f1, f2 = M.synth_features(31, 150, 48), M.synth_features(32, 130, 48, 0.4)
    got = float(metric.Fid_score.calculate_fid_from_features(torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda()))
    ref = M.frechet_distance(*M.feature_stats(f1), *M.feature_stats(f2))
    assert abs(got - ref) <= 1e-9 * max(1.0, abs(ref))
    with pytest.raises(RuntimeError, match="CUDA"):
        metric.Fid_score.calculate_activation_statistics_from_features(torch.zeros(4, 3))
    with pytest.raises(ValueError):
        metric.Fid_score.calculate_activation_statistics_from_features(torch.zeros(1, 3).cuda())
    with pytest.raises(AssertionError, match="different lengths"):
        z = torch.zeros(3, dtype=torch.float64).cuda()
        metric.Fid_score.calculate_frechet_distance(z, torch.eye(3).cuda(), torch.zeros(4).cuda(), torch.eye(4).cuda())


@pytest.mark.parametrize("i", range(len(M.METRIC_CASES["mmd"])))
def test_poly_mmd_matches_oracle(i):
    c, g = M.METRIC_CASES["mmd"][i], golden("metrics.npz")
    x = M.synth_features(c["seed"], c["m"], c["d"]).astype(np.float32)
    y = M.synth_features(c["seed"] + 100, c["m"], c["d"], c["shift"]).astype(np.float32)
    got = float(metric.kid.poly_mmd(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()))
    ref = float(g[f"mmd{i}"])
    assert abs(got - ref) <= 1e-9 * max(1.0, abs(ref)), (got, ref)


def test_kid_subsets_and_inception_score():
    g = golden("metrics.npz")
    x = torch.from_numpy(M.synth_features(41, 300, 32).astype(np.float32)).cuda()
    y = torch.from_numpy(M.synth_features(42, 280, 32, 0.5).astype(np.float32)).cuda()
    gen = torch.Generator().manual_seed(3)
    mean, std = metric.kid.kernel_inception_distance_from_features(x, y, subsets=5, subset_size=100, generator=gen)
    gen = torch.Generator().manual_seed(3)
    vals = []
    for _ in range(5):   # the same draws through the oracle
        pr, pf = torch.randperm(300, generator=gen)[:100], torch.randperm(280, generator=gen)[:100]
        vals.append(M.poly_mmd(x.cpu().numpy()[pr.numpy()], y.cpu().numpy()[pf.numpy()]))
    assert abs(float(mean) - np.mean(vals)) <= 1e-9 and abs(float(std) - np.std(vals)) <= 1e-9
    with pytest.raises(ValueError, match="subset_size"):
        metric.kid.kernel_inception_distance_from_features(x, y, subsets=1, subset_size=1000)
    for i, c in enumerate(M.METRIC_CASES["is"]):
        p = torch.from_numpy(M.synth_probs(c["seed"], c["N"], c["K"])).cuda()
        m, s = metric.IS_score.inception_score_from_probs(p, c["splits"])
        assert np.allclose([float(m), float(s)], g[f"is{i}"], rtol=1e-10, atol=1e-12)
    K = 10
    onehot = torch.eye(K, dtype=torch.float64)[torch.arange(100) % K].cuda()
    assert abs(float(metric.IS_score.inception_score_from_probs(onehot, 1)[0]) - K) < 1e-9


def test_recon_metrics_on_the_models_own_output():
    """The reconstruction test of R/main.py:303-323 end to end on the GPU: model forward -> MSE and 1 - SSIM, against the
    oracle's reconstruction pushed through the oracle's metrics."""
    from conftest import make_vqvae
    from oracle import snn_oracle as O
    from spiking_diffusion_b200 import synth
    from spiking_diffusion_b200.activation_based import functional
    T, B = 4, 8
    m, sd = make_vqvae(T, 128, seed=0)
    img = synth.synth_images(5, B)
    xs = img.unsqueeze(0).repeat(T, 1, 1, 1, 1)
    _, rec, _ = m(xs.cuda(), img.cuda())
    functional.reset_net(m)
    _, rec_ref, _ = O.vqvae_forward_eval(xs, sd)
    mse, ssim_loss = float(metric.mse_loss(rec, img.cuda())), 1 - float(metric.pytorch_ssim.SSIM(window_size=11)(rec, img.cuda()))
    assert abs(mse - M.mse(rec_ref, img)) <= 1e-6 and abs(ssim_loss - (1 - float(M.ssim(rec_ref, img)))) <= 1e-5
