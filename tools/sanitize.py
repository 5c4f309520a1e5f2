"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / synccheck) on the GPU box."""
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
torch.cuda.synchronize()
print("sanitize pass done", float(rec.abs().max()), int(tok.max()), float(pred.abs().max()), float(lg.abs().max()))
