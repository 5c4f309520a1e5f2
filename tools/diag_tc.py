"""Diagnostic for the tcgen05 conv kernel (run on the GPU box): controlled weight patterns isolate descriptor /
shift / split / epilogue problems.  Prints error statistics instead of asserting."""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import snn_oracle as O  # noqa: E402
from spiking_diffusion_b200 import _lib, engine  # noqa: E402
from spiking_diffusion_b200.activation_based import layer  # noqa: E402


def run(name, cin, cout, B, H, T, wfn, nsplit=2, rate=0.2):
    torch.manual_seed(0)
    conv = layer.Conv2d(cin, cout, 3, stride=1, padding=1)
    with torch.no_grad():
        conv.weight.zero_(); conv.bias.zero_()
        wfn(conv.weight)
    conv = conv.cuda().eval()
    try:
        fl = engine.FusedLayer(conv, None, None, T=T, B=B, H_in=H, W_in=H, in_kind=_lib.IN_STF,
                               out_kind=_lib.OUT_MEAN_T, impl="tc", nsplit=nsplit, in_T=1)
        s = (torch.rand(1, B, cin, H, H) < rate).float() * torch.randint(1, 4, (1, B, cin, H, H)).float()
        x = engine.stf_from_nchw(s.cuda())
        out = fl.run(x, fl.alloc_out())
        torch.cuda.synchronize()
        out = out.cpu()
        p = {"c.weight": conv.weight.detach().cpu(), "c.bias": conv.bias.detach().cpu()}
        ref = (O.conv_bn(s, p, "c", None, stride=1, padding=1)[0] / T).permute(0, 2, 3, 1)
        err = (out - ref).abs()
        print(f"[{name}] cin={cin} cout={cout} B={B} H={H} nsplit={nsplit}: max err {float(err.max()):.3e} "
              f"mean err {float(err.mean()):.3e} ref absmax {float(ref.abs().max()):.3e} "
              f"frac bad(>1e-3) {float((err > 1e-3).float().mean()):.4f}", flush=True)
        if float(err.max()) > 1e-3:
            bad = (err > 1e-3).nonzero()[:5]
            for b_ in bad:
                i = tuple(int(v) for v in b_)
                print("    at", i, "got", float(out[i]), "ref", float(ref[i]))
            # which rows / channels are bad?
            print("    bad per image:", (err > 1e-3).float().mean(dim=(1, 2, 3)).tolist()[:8])
            print("    bad per y:", (err > 1e-3).float().mean(dim=(0, 2, 3)).tolist())
            print("    bad per x:", (err > 1e-3).float().mean(dim=(0, 1, 3)).tolist())
            ch = (err > 1e-3).float().mean(dim=(0, 1, 2))
            print("    bad per channel (first 16):", ch[:16].tolist(), " any>0:", int((ch > 0).sum()), "of", cout)
    except Exception as e:  # noqa: BLE001
        print(f"[{name}] EXCEPTION {type(e).__name__}: {e}", flush=True)


def center_identity(w):
    n = min(w.shape[0], w.shape[1])
    for i in range(n):
        w[i, i, 1, 1] = 1.0


def tap_identity(ky, kx):
    def f(w):
        n = min(w.shape[0], w.shape[1])
        for i in range(n):
            w[i, i, ky, kx] = 1.0
    return f


def rand_w(w):
    w.copy_((torch.rand(w.shape) * 2 - 1) * 0.05)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    run("center identity, 1 K-block", 16, 32, 2, 7, 4, center_identity)
    run("center identity, 32ch", 32, 32, 2, 7, 4, center_identity)
    run("center identity, 64->128", 64, 128, 2, 7, 4, center_identity)
    for ky in range(3):
        for kx in range(3):
            run(f"tap({ky},{kx}) identity", 32, 32, 2, 7, 4, tap_identity(ky, kx))
    run("random nsplit=1", 64, 128, 2, 7, 4, rand_w, nsplit=1)
    run("random nsplit=2", 64, 128, 2, 7, 4, rand_w, nsplit=2)
    run("random many tiles", 128, 256, 40, 7, 4, rand_w)
    run("random 8x8 grid", 64, 128, 3, 8, 4, rand_w)
    print("diag done")
