"""Pin the oracle against the UNMODIFIED reference where it is mounted (the build container): fresh seeds, not the
committed fixtures.  Skipped on the GPU box, where /root/reference does not exist."""
import os

import pytest
import torch

from oracle import ref_loader, snn_oracle as O
from spiking_diffusion_b200 import synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference not mounted")


def test_vqvae_and_denoiser_bit_identical_at_T16():
    R = ref_loader.load()
    T = 16
    sd = synth.synth_vqvae_state(7, T=T)
    m = R.SNN_VQVAE(1, 16, 128, torch.tensor(1.0))
    R.functional.set_step_mode(m, "m")
    m.load_state_dict(sd)
    m.eval()
    img = synth.synth_images(7, 3)
    xs = img.unsqueeze(0).repeat(T, 1, 1, 1, 1)
    with torch.inference_mode():
        e, rec, idx = m(xs, img)
    R.functional.reset_net(m)
    e_o, rec_o, idx_o = O.vqvae_forward_eval(xs, sd)
    assert torch.equal(e, e_o) and torch.equal(rec, rec_o) and torch.equal(idx, idx_o)
    dsd = synth.synth_denoiser_state(7)
    d = R.DummyModel(1, 128)
    R.functional.set_step_mode(d, "m")
    d.load_state_dict(dsd)
    d.eval()
    x = torch.randint(0, 129, (2, 1, 7, 7)).float()
    t = torch.randint(1, 50, (2,))
    with torch.inference_mode():
        lg = d(x, t)
    assert torch.equal(lg, O.denoiser_forward(x, t, dsd, T))


@pytest.mark.parametrize("vr,detach,decay", [(0.0, False, True), (None, False, True), (-0.5, True, True), (0.0, False, False)])
def test_training_branch_gradients_match_reference(vr, detach, decay):
    """Surrogate-gradient BPTT: the oracle's autograd restatement vs the reference's LIFNode in train mode."""
    R = ref_loader.load()
    g = torch.Generator().manual_seed(3)
    x = ((torch.rand(5, 4, 9, generator=g) - 0.3) * 3).requires_grad_(True)
    w = torch.rand(5, 4, 9, generator=g)
    n = R.neuron.LIFNode(tau=2.0, decay_input=decay, v_threshold=1.0, v_reset=vr, surrogate_function=R.surrogate.ATan(),
                         detach_reset=detach, step_mode="m").train()
    s = n(x)
    ((s * w).sum() + (n.v * 0.3).sum()).backward()
    x2 = x.detach().clone().requires_grad_(True)
    s2, v2 = O.lif_multi_step_train(x2, None, 2.0, 1.0, vr, decay, detach, 2.0)
    ((s2 * w).sum() + (v2 * 0.3).sum()).backward()
    assert torch.equal(s, s2)
    assert torch.allclose(x.grad, x2.grad, rtol=1e-6, atol=1e-7)
    assert float(x.grad.abs().max()) > 0


def _leaf(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "coef" not in k else v.clone())
            for k, v in sd.items()}


def test_vqvae_training_step_matches_reference():
    """Losses and parameter gradients of one training forward/backward (R/main.py:131-141) at T=16."""
    R = ref_loader.load()
    T, B = 16, 3
    sd = synth.synth_vqvae_state(5, T=T)
    m = R.SNN_VQVAE(1, 16, 128, torch.tensor(0.09))
    R.functional.set_step_mode(m, "m")
    m.load_state_dict(sd)
    m.train()
    img = synth.synth_images(5, B)
    xs = img.unsqueeze(0).repeat(T, 1, 1, 1, 1)
    e_q, rec, real = m(xs, img)
    (e_q + rec).backward()
    p = _leaf(sd)
    e_q2, rec2, real2 = O.vqvae_forward_train(xs, img, p, torch.tensor(0.09))
    (e_q2 + rec2).backward()
    assert torch.allclose(e_q, e_q2, rtol=1e-6, atol=1e-7) and torch.allclose(rec, rec2, rtol=1e-6, atol=1e-7)
    ref = dict(m.named_parameters())
    for k in ("encoder.snn_convs.0.weight", "encoder.snn_convs.4.weight", "vq_layer.alpha", "vq_layer.embeddings.weight",
              "vq_layer.poisson.0.weight", "decoder.snn_convs.3.weight", "decoder.snn_convs.6.bias"):
        assert torch.allclose(ref[k].grad, p[k].grad, rtol=1e-5, atol=1e-8), k
    assert torch.allclose(m.encoder.snn_convs[1].running_var, p["encoder.snn_convs.1.running_var"], rtol=1e-6)


def test_denoiser_training_loss_matches_reference():
    R = ref_loader.load()
    T, b, K = 16, 2, 128
    sd = synth.synth_denoiser_state(5, num_embeddings=K)
    d = R.DummyModel(1, K)
    R.functional.set_step_mode(d, "m")
    d.load_state_dict(sd)
    d.train()
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, K + 1, (b, 1, 7, 7), generator=g).float()
    t = torch.randint(1, 50, (b,), generator=g)
    lg = d(x, t)
    p = _leaf(sd)
    lg2 = O.denoiser_forward_train(x, t, p, T)
    assert torch.allclose(lg, lg2, rtol=1e-5, atol=1e-6)
    tgt = torch.randint(0, K, (b, 1, 7, 7), generator=g)
    tgt[x.long() != K] = -1
    l1 = O.diffusion_train_loss(lg, tgt, t, 49)
    l2 = O.diffusion_train_loss(lg2, tgt, t, 49)
    l1.backward(); l2.backward()
    ref = dict(d.named_parameters())
    for k in ("conv1.0.weight", "conv4.0.weight", "conv6.0.bias", "conv3.1.weight"):
        assert torch.allclose(ref[k].grad, p[k].grad, rtol=1e-4, atol=1e-8), k
