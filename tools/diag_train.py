import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_vqvae
from oracle import snn_oracle as O
from spiking_diffusion_b200 import synth
from spiking_diffusion_b200.activation_based import neuron

T, B, K = 4, 4, 128
m, sd = make_vqvae(T, K, seed=6)
m.data_variance = torch.tensor(0.09)
m.train()
img = synth.synth_images(6, B)
xs = img.unsqueeze(0).repeat(T, 1, 1, 1, 1)
p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "coef" not in k else v.clone()) for k, v in sd.items()}
# oracle with spike recording
rec_ref = []
orig = O.lif_multi_step_train
def rec_lif(*a, **k):
    s, v = orig(*a, **k); rec_ref.append(s.detach()); return s, v
O.lif_multi_step_train = rec_lif
e_ref, r_ref, _ = O.vqvae_forward_train(xs, img, p, torch.tensor(0.09))
(e_ref + r_ref).backward()
outs = []
for mod in m.modules():
    if isinstance(mod, neuron.LIFNode):
        mod.register_forward_hook(lambda mo, i, o: outs.append(o.detach().cpu()))
e_q, rec, _ = m(xs.cuda(), img.cuda())
(e_q + rec).backward()
print("losses", float(e_q), float(e_ref), float(rec), float(r_ref))
for i, (a, b) in enumerate(zip(outs, rec_ref)):
    print("lif", i, "flips", int((a != b).sum()), "of", a.numel())
named = dict(m.named_parameters())
big = max(float(p[k].grad.norm()) for k in named)
for k, v in named.items():
    a, b = v.grad.detach().cpu().double(), p[k].grad.double()
    print(f"{k:36s} |g_ref| {float(b.norm()):.3e}  rel err {float((a-b).norm()/max(float(b.norm()),1e-4*big)):.2e}")
