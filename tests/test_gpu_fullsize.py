"""Full-size checks at BASELINE.json's shapes through size-independent properties (the CPU oracle would take minutes
here): determinism, graph replay == eager launch, shard union == whole batch, no mask tokens left, linearity of the
linear read-out, decode idempotence, and STF round trips."""
import pytest
import torch

from conftest import make_denoiser, make_vqvae
from spiking_diffusion_b200 import _lib, engine
from spiking_diffusion_b200.snn_model.vq_diffusion import AbsorbingDiffusion

pytestmark = pytest.mark.gpu


def test_cfg2_sampling_properties_b256():
    T, K, b = 4, 128, 256
    den, _ = make_denoiser(T, K, seed=0)
    ab = AbsorbingDiffusion(den, mask_id=K, shape=(7, 7), n_samples=b)
    plan = ab.plan(b)
    x_graph = plan.sample(0.65, 49, seed=11, use_graph=True).clone()
    x_eager = plan.sample(0.65, 49, seed=11, use_graph=False).clone()
    x_again = plan.sample(0.65, 49, seed=11, use_graph=True).clone()
    x_other = plan.sample(0.65, 49, seed=12, use_graph=True).clone()
    assert torch.equal(x_graph, x_eager) and torch.equal(x_graph, x_again)     # replay == eager, deterministic
    assert not torch.equal(x_graph, x_other)                                   # the seed reaches the device RNG state
    assert int(x_graph.max()) < K and int(x_graph.min()) >= 0                  # fully unmasked
    assert x_graph.unique().numel() > K // 2                                   # a diverse sample, not a collapsed one
    # shard union == whole batch, at full size, 4 ragged shards
    parts, lo = [], 0
    for n in (100, 28, 64, 64):
        a = AbsorbingDiffusion(den, mask_id=K, shape=(7, 7), n_samples=n)
        parts.append(a.sample(temp=0.65, sample_steps=49, seed=11, n_global=b, shard_base=lo))
        lo += n
    assert torch.equal(torch.cat(parts), x_graph)
    # fewer steps than tokens is "skipping": still terminates fully unmasked because 1/t reaches 1 at t = 1
    x_short = ab.sample(temp=1.0, sample_steps=10, seed=3)
    assert int(x_short.max()) < K


def test_cfg3_cifar_shape_pipeline_b64():
    """8x8 latent, 3-channel 32x32 images (the extrapolated CIFAR-10 config, SURVEY.md finding 2)."""
    from spiking_diffusion_b200 import synth
    from spiking_diffusion_b200.activation_based import functional
    from spiking_diffusion_b200.snn_model import SNN_VQVAE, DummyModel
    T, K, b = 4, 128, 64
    vae = SNN_VQVAE(3, 16, K, torch.tensor(1.0), T=T)
    functional.set_step_mode(vae, "m")
    vae.load_state_dict(synth.synth_vqvae_state(0, in_dim=3, num_embeddings=K, T=T))
    vae = vae.eval().cuda()
    den = DummyModel(1, K, T=T)
    functional.set_step_mode(den, "m")
    den.load_state_dict(synth.synth_denoiser_state(0, num_embeddings=K, num_timesteps=64))
    den = den.eval().cuda()
    ab = AbsorbingDiffusion(den, mask_id=K, shape=(8, 8), n_samples=b)
    tok = ab.sample(temp=1.0, sample_steps=64, seed=1)
    assert tok.shape == (b, 1, 8, 8) and int(tok.max()) < K
    pred = vae.decode_indices(tok.reshape(b, 8, 8))
    assert pred.shape == (b, 3, 32, 32) and float(pred.abs().max()) <= 1.0
    assert torch.equal(pred, vae.decode_indices(tok.reshape(b, 8, 8)))         # decode is a pure function of the tokens
    img = torch.rand(b, 3, 32, 32, device="cuda") - 0.5
    e, rec, idx = vae(img.unsqueeze(0).repeat(T, 1, 1, 1, 1), img)
    assert rec.shape == (b, 3, 32, 32) and idx.shape == (b * 64,) and e.shape == (T, b, 16, 8, 8)


def test_cfg4_T8_K512_denoiser_linear_readout_b64():
    """conv6 is linear in the T-summed spikes: logits(2*counts) - bias == 2*(logits(counts) - bias)."""
    T, K, b = 8, 512, 64
    den, sd = make_denoiser(T, K, seed=1)
    dp = den.plan(b, 7, 7)
    x = torch.full((b, 1, 7, 7), float(K), device="cuda")
    x[::2] = 3.0
    t = torch.randint(1, 50, (b,), device="cuda")
    lg = dp.run(x, t).clone()
    bias = sd["conv6.0.bias"].cuda()
    x5s, x1s = dp.x5s.clone(), dp.x1s.clone()
    lg2 = dp.l6.run((x5s * 2).contiguous(), dp.l6.alloc_out(), x2=(x1s * 2).contiguous())
    assert float(((lg2 - bias) - 2 * (lg - bias)).abs().max()) <= 2e-4
    assert lg.shape == (b, 7, 7, K) and bool(torch.isfinite(lg).all())


@pytest.mark.parametrize("shape", [(4, 256, 512, 7, 7), (8, 3, 24, 5, 9), (1, 1, 1, 1, 1)])
def test_stf_round_trip(shape):
    x = (torch.rand(shape, device="cuda") < 0.1).float()
    T, B, C, H, W = shape
    stf = engine.stf_from_nchw(x)
    assert torch.equal(engine.stf_to_nchw(stf, T, B, C, H, W), x)
    assert float(stf.float().sum()) == float(x.sum())          # guard rows / padded channels are zero


@pytest.mark.parametrize("T", [8, 16])
def test_multi_pass_shards_equal_the_whole_batch_across_pass_modes(T):
    """T = 8 / 16 layers run as fused passes or, for small shards, as T-parallel passes plus a separate LIF kernel; which
    one is chosen depends on the shard size and the layer.  Ragged shards must still reproduce the whole batch bit for
    bit (same K order, same LIF arithmetic)."""
    K, b = 128, 100
    den, _ = make_denoiser(T, K, seed=4)
    whole = AbsorbingDiffusion(den, mask_id=K, shape=(7, 7), n_samples=b).sample(temp=0.8, sample_steps=20, seed=21)
    parts, lo = [], 0
    for n in (7, 33, 60):
        a = AbsorbingDiffusion(den, mask_id=K, shape=(7, 7), n_samples=n)
        parts.append(a.sample(temp=0.8, sample_steps=20, seed=21, n_global=b, shard_base=lo))
        lo += n
    assert torch.equal(torch.cat(parts), whole)
    assert int(whole.max()) < K
