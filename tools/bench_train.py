"""Training-iteration timing (first-correct-version kernels) vs the oracle's CPU autograd, reference batch size 32."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_denoiser, make_vqvae  # noqa: E402
from oracle import snn_oracle as O  # noqa: E402
from spiking_diffusion_b200 import synth  # noqa: E402
from spiking_diffusion_b200.activation_based import functional  # noqa: E402


def gpu_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def main():
    out = {}
    for T in (4, 16):
        B = 32
        m, sd = make_vqvae(T, 128, seed=0)
        m.data_variance = torch.tensor(0.09)
        m.train()
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=0.001)
        img = synth.synth_images(0, B).cuda()
        xs = img.unsqueeze(0).repeat(T, 1, 1, 1, 1)

        def step():
            e_q, rec, _ = m(xs, img)
            opt.zero_grad(); (e_q + rec).backward(); opt.step(); functional.reset_net(m)
        ms = gpu_time(step)
        p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "coef" not in k else v.clone())
             for k, v in sd.items()}
        torch.set_num_threads(os.cpu_count())
        t0 = time.perf_counter()
        e, r, _ = O.vqvae_forward_train(xs.cpu(), img.cpu(), p, torch.tensor(0.09)); (e + r).backward()
        cpu_ms = (time.perf_counter() - t0) * 1e3
        out[f"vqvae train iter B=32 T={T}"] = {"gpu_ms": round(ms, 2), "cpu_oracle_ms": round(cpu_ms, 1), "images_per_s": round(B / ms * 1e3, 1)}
        d, dsd = make_denoiser(T, 128, seed=0)
        d.train()
        opt2 = torch.optim.AdamW(d.parameters(), lr=1e-3, weight_decay=0.001)
        x = torch.randint(0, 129, (B, 1, 7, 7)).float().cuda()
        t = torch.randint(1, 50, (B,)).cuda()
        tgt = torch.randint(0, 128, (B, 1, 7, 7)).cuda()

        def dstep():
            lg = d(x, t)
            loss = O.diffusion_train_loss(lg, tgt, t, 49)
            opt2.zero_grad(); loss.backward(); opt2.step(); functional.reset_net(d)
        ms = gpu_time(dstep, reps=3)
        out[f"denoiser train iter b=32 T={T}"] = {"gpu_ms": round(ms, 2), "images_per_s": round(B / ms * 1e3, 1)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
