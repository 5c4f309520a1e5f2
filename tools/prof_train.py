"""Two denoiser training iterations (b=32, T=4) for an ncu launch list.  Usage: ncu ... python tools/prof_train.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_denoiser  # noqa: E402
from oracle import snn_oracle as O  # noqa: E402
from spiking_diffusion_b200.activation_based import functional  # noqa: E402

T, B = int(os.environ.get("T", "4")), 32
d, dsd = make_denoiser(T, 128, seed=0)
d.train()
opt = torch.optim.AdamW(d.parameters(), lr=1e-3, weight_decay=0.001)
x = torch.randint(0, 129, (B, 1, 7, 7)).float().cuda()
t = torch.randint(1, 50, (B,)).cuda()
tgt = torch.randint(0, 128, (B, 1, 7, 7)).cuda()
for _ in range(2):
    lg = d(x, t)
    loss = O.diffusion_train_loss(lg, tgt, t, 49)
    opt.zero_grad(); loss.backward(); opt.step(); functional.reset_net(d)
torch.cuda.synchronize()
print("loss", float(loss))
