"""Random-configuration check of the training-path convolution kernels (tiled forward / input gradient, split-reduction
weight gradient) against torch's float64 convolutions on the GPU.  Usage: python tools/fuzz_train_conv.py [n] [seed]"""
import os
import random
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spiking_diffusion_b200.activation_based import layer  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp(min=1e-20))


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad, worst = 0, 0.0
    for case in range(n_cases):
        transposed = rng.random() < 0.4
        k = rng.choice([1, 3, 3, 4, 5])
        stride = rng.choice([1, 1, 2, 3])
        pad = rng.randrange(0, k // 2 + 1) if k > 1 else 0
        cin, cout = rng.choice([1, 2, 3, 16, 33, 64, 130]), rng.choice([1, 3, 16, 40, 64, 129])
        H, W = rng.choice([3, 5, 7, 8, 14, 17]), rng.choice([3, 5, 7, 8, 14, 17])
        T, B = rng.choice([1, 2, 4]), rng.choice([1, 2, 5])
        op = rng.randrange(0, stride) if transposed else 0
        if (H + 2 * pad < k or W + 2 * pad < k) and not transposed:
            continue
        g = torch.Generator().manual_seed(case)
        if transposed:
            m = layer.ConvTranspose2d(cin, cout, k, stride=stride, padding=pad, output_padding=op, step_mode="m")
        else:
            m = layer.Conv2d(cin, cout, k, stride=stride, padding=pad, step_mode="m")
        m = m.cuda()
        x = torch.randn(T, B, cin, H, W, generator=g).cuda().requires_grad_(True)
        y = m(x)
        gy = torch.randn(y.shape, generator=g).cuda()
        y.backward(gy)
        # reference in float64 (cuDNN's fp32 algorithms - Winograd / FFT for 5x5 stride 1 - are themselves only good to ~1e-4)
        xr = x.detach().double().requires_grad_(True)
        w, b = m.weight.detach().double().requires_grad_(True), m.bias.detach().double().requires_grad_(True)
        if transposed:
            yr = F.conv_transpose2d(xr.flatten(0, 1), w, b, stride=stride, padding=pad, output_padding=op)
        else:
            yr = F.conv2d(xr.flatten(0, 1), w, b, stride=stride, padding=pad)
        yr.backward(gy.flatten(0, 1).double())
        errs = dict(y=rel(y.flatten(0, 1).double(), yr.detach()), gx=rel(x.grad.double(), xr.grad),
                    gw=rel(m.weight.grad.double(), w.grad), gb=rel(m.bias.grad.double(), b.grad))
        w_ = max(errs.values())
        worst = max(worst, w_)
        ok = w_ <= 1e-5
        bad += not ok
        print(f"case {case}: {'convT' if transposed else 'conv '} {cin}->{cout} k{k} s{stride} p{pad} op{op} {H}x{W} T={T} B={B}: "
              + " ".join(f"{n} {v:.1e}" for n, v in errs.items()), "" if ok else "<-- CHECK", flush=True)
    print(f"worst relative error {worst:.2e}; suspicious cases: {bad}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
