"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / synccheck) on the GPU box: every kernel family
of the library is launched at least once."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SD_SAMPLER_GRAPH", "0")
from spiking_diffusion_b200 import engine, synth  # noqa: E402
from spiking_diffusion_b200.activation_based import functional, neuron  # noqa: E402
from spiking_diffusion_b200.snn_model import SNN_VQVAE, DummyModel, AbsorbingDiffusion  # noqa: E402

T, K, B = 4, 128, 3
vae = SNN_VQVAE(1, 16, K, torch.tensor(1.0), T=T)
den = DummyModel(1, K, T=T)
functional.set_step_mode(vae, "m"); functional.set_step_mode(den, "m")
vae.load_state_dict(synth.synth_vqvae_state(0, T=T)); den.load_state_dict(synth.synth_denoiser_state(0))
vae, den = vae.eval().cuda(), den.eval().cuda()
img = synth.synth_images(0, B).cuda()
e, rec, idx = vae(img.unsqueeze(0).repeat(T, 1, 1, 1, 1), img)
functional.reset_net(vae)
ab = AbsorbingDiffusion(den, mask_id=K, shape=(7, 7), n_samples=B)
tok = ab.sample(temp=1.0, sample_steps=3, seed=0)
pred = vae.decode_indices(tok.reshape(B, 7, 7))
n = neuron.LIFNode(step_mode="m").eval()
n(torch.rand(4, 1001, device="cuda"))
den8 = DummyModel(1, K, T=8)
functional.set_step_mode(den8, "m")
den8.load_state_dict(synth.synth_denoiser_state(0))
den8 = den8.eval().cuda()
lg = den8(torch.full((2, 1, 7, 7), float(K), device="cuda"), torch.ones(2, dtype=torch.long, device="cuda"))
# CTA-pair kernel with two concurrent sub-batch streams (70 images -> 36 + 34), fused VQ-VAE plan (tensor-core
# encoder / decoder routes), and one training iteration of each model (tiled conv / weight-gradient / BN / LIF backward)
ab70 = AbsorbingDiffusion(den, mask_id=K, shape=(7, 7), n_samples=70)
tok70 = ab70.sample(temp=1.0, sample_steps=2, seed=1)
plan = vae.plan(T, 5, 28, 28)
plan.forward(synth.synth_images(1, 5).cuda(), const_over_T=True)
vae.train(); den.train()
vae.data_variance = torch.tensor(0.09)
img2 = synth.synth_images(2, 4).cuda()
e_q, r_l, _ = vae(img2.unsqueeze(0).repeat(T, 1, 1, 1, 1), img2)
(e_q + r_l).backward()
functional.reset_net(vae)
lg_t = den(torch.randint(0, K + 1, (4, 1, 7, 7)).float().cuda(), torch.randint(1, 50, (4,)).cuda())
lg_t.square().mean().backward()
functional.reset_net(den)
# round 2: the kind::f16 denoiser path (the default is kind::i8, exercised above), a T-parallel int8 layer (lone small
# batch at T = 8 -> above), chunked sampling from one plan, and the metric kernels
den.eval(); vae.eval()
den.nsplit = 2
ab2 = AbsorbingDiffusion(den, mask_id=K, shape=(7, 7), n_samples=5)
ab2.max_plan_batch = 2
tok5 = ab2.sample(temp=1.0, sample_steps=2, seed=3)
den.nsplit = 3
from spiking_diffusion_b200 import metric  # noqa: E402
a_img, b_img = torch.rand(3, 1, 28, 28, device="cuda") - 0.5, torch.rand(3, 1, 28, 28, device="cuda") - 0.5
ssim_v, mse_v = metric.pytorch_ssim.SSIM(window_size=11)(a_img, b_img), metric.mse_loss(a_img, b_img)
f1, f2 = torch.randn(40, 33, device="cuda", dtype=torch.float64), torch.randn(50, 33, device="cuda", dtype=torch.float64) + 0.1
fid = metric.Fid_score.calculate_fid_from_features(f1, f2)
mmd = metric.kid.poly_mmd(f1[:37].float(), f2[:37].float())
is_m, is_s = metric.IS_score.inception_score_from_probs(torch.softmax(torch.randn(30, 17, device="cuda", dtype=torch.float64), 1), 3)
torch.cuda.synchronize()
print("metrics", float(ssim_v), float(mse_v), float(fid), float(mmd), float(is_m))
print("sanitize pass done", float(rec.abs().max()), int(tok.max()), float(pred.abs().max()), float(lg.abs().max()))
