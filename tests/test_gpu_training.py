"""(f) rank 1, training path on the GPU: conv forward / input-gradient / weight-gradient, train-mode BatchNorm and the
model-level training branches against torch autograd through the CPU oracle (which tests/test_oracle_pin.py pins to
the reference's own train-mode losses and gradients).  Tolerance: fp32, relative L2 error <= 1e-4 for gradients
(elementwise comparison is meaningless where a near-threshold spike differs), losses rtol 1e-4."""
import pytest
import torch
import torch.nn.functional as F

from conftest import make_denoiser, make_vqvae
from oracle import snn_oracle as O
from spiking_diffusion_b200 import synth
from spiking_diffusion_b200.activation_based import functional, layer

pytestmark = pytest.mark.gpu


def rel_err(a, b, floor=1e-30):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(floor))


def worst_grad_err(named, p):
    """Largest relative L2 gradient error over all parameters.  A conv bias that feeds a train-mode BatchNorm has an
    exactly-zero gradient (the batch mean absorbs it), so both sides hold rounding noise there: errors are measured
    against max(|g_ref|, 1e-2 * largest gradient norm of the model); measured: every non-degenerate gradient agrees to
    better than 1e-6 relative (tools/diag_train.py)."""
    big = max(float(p[k].grad.norm()) for k in named)
    return max(rel_err(v.grad, p[k].grad, floor=1e-2 * big) for k, v in named.items())


@pytest.mark.parametrize("cfg", [
    dict(cin=3, cout=8, k=3, s=1, p=1, H=9), dict(cin=8, cout=16, k=3, s=2, p=1, H=14), dict(cin=16, cout=4, k=1, s=1, p=0, H=7),
    dict(cin=6, cout=10, k=3, s=2, p=1, H=7, tr=True, op=1), dict(cin=10, cout=3, k=3, s=1, p=1, H=8, tr=True, op=0),
])
def test_conv_forward_and_gradients(cfg):
    T, B, H = 3, 2, cfg["H"]
    g = torch.Generator().manual_seed(1)
    if cfg.get("tr"):
        m = layer.ConvTranspose2d(cfg["cin"], cfg["cout"], cfg["k"], stride=cfg["s"], padding=cfg["p"], output_padding=cfg["op"],
                                  step_mode="m")
        fn = lambda x, w, b: F.conv_transpose2d(x, w, b, stride=cfg["s"], padding=cfg["p"], output_padding=cfg["op"])
    else:
        m = layer.Conv2d(cfg["cin"], cfg["cout"], cfg["k"], stride=cfg["s"], padding=cfg["p"], step_mode="m")
        fn = lambda x, w, b: F.conv2d(x, w, b, stride=cfg["s"], padding=cfg["p"])
    x = torch.randn(T, B, cfg["cin"], H, H, generator=g)
    w_ref = m.weight.detach().clone().requires_grad_(True)
    b_ref = m.bias.detach().clone().requires_grad_(True)
    x_ref = x.clone().requires_grad_(True)
    y_ref = fn(x_ref.flatten(0, 1), w_ref, b_ref)
    gy = torch.randn(y_ref.shape, generator=g)
    (y_ref * gy).sum().backward()
    m = m.cuda().train()
    xg = x.cuda().requires_grad_(True)
    y = m(xg)
    assert y.shape == (T, B) + y_ref.shape[1:]
    (y * gy.view(y.shape).cuda()).sum().backward()
    assert rel_err(y, y_ref.view(y.shape)) <= 1e-5
    assert rel_err(xg.grad, x_ref.grad) <= 1e-5
    assert rel_err(m.weight.grad, w_ref.grad) <= 1e-5
    assert rel_err(m.bias.grad, b_ref.grad) <= 1e-5


@pytest.mark.parametrize("cfg", [dict(cin=64, cout=128, T=4, B=3, H=7), dict(cin=32, cout=16, T=2, B=5, H=8),
                                 dict(cin=128, cout=48, T=8, B=2, H=5)])
def test_spike_input_conv_forward_runs_on_tensor_cores(cfg, monkeypatch):
    """Training branch: a 3x3 stride-1 layer.Conv2d fed by a LIFNode's spikes takes the tcgen05 kind::i8 kernel for its
    forward (three exact int8 weight digits: 22-bit weights); output and all gradients against torch autograd, same bars
    as the CUDA-core path.  SD_TRAIN_TC=0 switches it off."""
    T, B, H, cin, cout = cfg["T"], cfg["B"], cfg["H"], cfg["cin"], cfg["cout"]
    g = torch.Generator().manual_seed(3)
    m = layer.Conv2d(cin, cout, 3, stride=1, padding=1, step_mode="m")
    x = (torch.rand(T, B, cin, H, H, generator=g) < 0.2).float()
    w_ref = m.weight.detach().clone().requires_grad_(True)
    b_ref = m.bias.detach().clone().requires_grad_(True)
    x_ref = x.clone().requires_grad_(True)
    y_ref = F.conv2d(x_ref.flatten(0, 1), w_ref, b_ref, stride=1, padding=1)
    gy = torch.randn(y_ref.shape, generator=g)
    (y_ref * gy).sum().backward()
    m = m.cuda().train()
    xg = x.cuda().requires_grad_(True)
    xg._sd_is_spikes = True                      # what LIFNode's training branch puts on its output
    y_tc = layer._ConvFn._tc_forward(xg.detach(), m)
    assert y_tc is not None and rel_err(y_tc, y_ref.view(y_tc.shape)) <= 1e-6      # the tensor-core result itself
    y = m(xg)
    assert torch.equal(y.detach(), y_tc)         # ... and the module took that path
    (y * gy.view(y.shape).cuda()).sum().backward()
    assert rel_err(xg.grad, x_ref.grad) <= 1e-5
    assert rel_err(m.weight.grad, w_ref.grad) <= 1e-5
    assert rel_err(m.bias.grad, b_ref.grad) <= 1e-5
    monkeypatch.setenv("SD_TRAIN_TC", "0")
    assert layer._ConvFn._tc_forward(xg.detach(), m) is None
    # an untagged input or a shape the kernel does not take stays on the CUDA-core kernel
    m2 = layer.Conv2d(24, 16, 3, stride=1, padding=1, step_mode="m").cuda().train()
    monkeypatch.delenv("SD_TRAIN_TC")
    assert layer._ConvFn._tc_forward(torch.zeros(T, B, 24, H, H, device="cuda"), m2) is None


def test_batchnorm_train_mode_forward_backward_and_running_stats():
    C = 12
    g = torch.Generator().manual_seed(2)
    x = torch.randn(4, 3, C, 5, 6, generator=g) * 2 + 0.5
    ref = torch.nn.BatchNorm2d(C)
    bn = layer.BatchNorm2d(C, step_mode="m")
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5, generator=g); ref.bias.uniform_(-1, 1, generator=g)
    bn.load_state_dict(ref.state_dict())
    x_ref = x.clone().requires_grad_(True)
    y_ref = ref.train()(x_ref.flatten(0, 1))
    gy = torch.randn(y_ref.shape, generator=g)
    (y_ref * gy).sum().backward()
    bn = bn.cuda().train()
    xg = x.cuda().requires_grad_(True)
    y = bn(xg)
    (y * gy.view(y.shape).cuda()).sum().backward()
    assert rel_err(y, y_ref.view(y.shape)) <= 1e-5 and rel_err(xg.grad, x_ref.grad.view(x.shape)) <= 1e-4
    assert rel_err(bn.weight.grad, ref.weight.grad) <= 1e-5 and rel_err(bn.bias.grad, ref.bias.grad) <= 1e-5
    assert rel_err(bn.running_mean, ref.running_mean) <= 1e-5 and rel_err(bn.running_var, ref.running_var) <= 1e-5
    assert int(bn.num_batches_tracked) == 1


def _leaf(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "coef" not in k else v.clone())
            for k, v in sd.items()}


def test_vqvae_training_step_vs_oracle():
    """One forward/backward of R/main.py:131-141 (T=4): losses and gradients of every parameter group."""
    T, B, K = 4, 4, 128
    m, sd = make_vqvae(T, K, seed=6)
    m.data_variance = torch.tensor(0.09)
    m.train()
    img = synth.synth_images(6, B)
    xs = img.unsqueeze(0).repeat(T, 1, 1, 1, 1)
    p = _leaf(sd)
    e_ref, r_ref, real_ref = O.vqvae_forward_train(xs, img, p, torch.tensor(0.09))
    (e_ref + r_ref).backward()
    e_q, rec, real = m(xs.cuda(), img.cuda())
    (e_q + rec).backward()
    functional.reset_net(m)
    assert torch.allclose(e_q.cpu(), e_ref, rtol=1e-4) and torch.allclose(rec.cpu(), r_ref, rtol=1e-4)
    named = dict(m.named_parameters())
    assert all(v.grad is not None for v in named.values())
    worst = worst_grad_err(named, p)
    assert worst <= 1e-3, worst
    assert rel_err(m.encoder.snn_convs[1].running_mean, p["encoder.snn_convs.1.running_mean"]) <= 1e-4
    # an optimiser step runs on these modules as on the reference's (R/main.py:113-116)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=0.001)
    opt.step()


def test_denoiser_training_loss_vs_oracle():
    from spiking_diffusion_b200.snn_model import AbsorbingDiffusion
    T, K, b = 4, 128, 3
    m, sd = make_denoiser(T, K, seed=7)
    m.train()
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, K + 1, (b, 1, 7, 7), generator=g).float()
    t = torch.randint(1, 50, (b,), generator=g)
    tgt = torch.randint(0, K, (b, 1, 7, 7), generator=g)
    tgt[x.long() != K] = -1
    p = _leaf(sd)
    lg_ref = O.denoiser_forward_train(x, t, p, T)
    l_ref = O.diffusion_train_loss(lg_ref, tgt, t, 49)
    l_ref.backward()
    lg = m(x.cuda(), t.cuda())
    loss = O.diffusion_train_loss(lg, tgt.cuda(), t.cuda(), 49)
    loss.backward()
    functional.reset_net(m)
    assert rel_err(lg, lg_ref) <= 1e-4 and torch.allclose(loss.cpu(), l_ref, rtol=1e-4)
    worst = worst_grad_err(dict(m.named_parameters()), p)
    assert worst <= 1e-3, worst
    # the reference's own training entry point runs end to end (vq_diffusion.py:75-101, 144-147)
    ab = AbsorbingDiffusion(m, mask_id=K)
    stats = ab.train_iter(torch.randint(0, K, (b, 1, 7, 7)).float().cuda())
    stats["loss"].backward()
    assert bool(torch.isfinite(stats["loss"]))
