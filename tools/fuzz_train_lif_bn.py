"""Random-shape check of the training-mode LIF neuron (surrogate-gradient BPTT kernel) against the oracle's autograd and
of train-mode BatchNorm (forward, backward, running statistics) against torch in float64.
Usage: python tools/fuzz_train_lif_bn.py [n] [seed]"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import snn_oracle as O  # noqa: E402  (checker)
from spiking_diffusion_b200.activation_based import layer, neuron, surrogate  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-20))


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    for case in range(n_cases):
        g = torch.Generator().manual_seed(case)
        T = rng.choice([1, 2, 3, 4, 8, 16])
        shape = rng.choice([(1, 1, 1, 1), (2, 3, 5, 7), (3, 16, 7, 7), (1, 5, 1, 9), (4, 33, 3, 3), (2, 64, 14, 14)])
        tau, v_th = rng.choice([2.0, 2.0, 3.0, 1.5]), rng.choice([1.0, 0.5])
        v_reset = rng.choice([0.0, 0.0, None, -0.2])
        decay, detach = rng.random() < 0.8, rng.random() < 0.3
        # ---- LIF ----
        x = (torch.randn((T,) + shape, generator=g) * 1.2 + 0.4)
        n = neuron.LIFNode(tau=tau, decay_input=decay, v_threshold=v_th, v_reset=v_reset, detach_reset=detach,
                           surrogate_function=surrogate.ATan(), step_mode="m").train()
        xg = x.cuda().requires_grad_(True)
        s = n(xg)
        gs = torch.randn(s.shape, generator=g)
        s.backward(gs.cuda())
        xr = x.clone().requires_grad_(True)
        s_ref, v_ref = O.lif_multi_step_train(xr, tau=tau, v_threshold=v_th, v_reset=v_reset, decay_input=decay,
                                              detach_reset=detach)
        s_ref.backward(gs)
        same = torch.equal(s.detach().cpu(), s_ref.detach())
        e_lif = rel(xg.grad.cpu(), xr.grad) if same else float("nan")
        e_v = rel(n.v.cpu(), v_ref.detach()) if same and float(v_ref.abs().max()) > 0 else 0.0
        # ---- BatchNorm, train mode ----
        C = shape[1]
        xb = torch.randn((T,) + shape, generator=g) * rng.choice([0.5, 3.0]) + rng.choice([0.0, 2.0])
        bn = layer.BatchNorm2d(C, step_mode="m").cuda().train()
        ref = torch.nn.BatchNorm2d(C).double().train()
        with torch.no_grad():
            w0, b0 = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
            bn.weight.copy_(w0); bn.bias.copy_(b0); ref.weight.copy_(w0); ref.bias.copy_(b0)
        xbg = xb.cuda().requires_grad_(True)
        yb = bn(xbg)
        gyb = torch.randn(yb.shape, generator=g)
        yb.backward(gyb.cuda())
        xbr = xb.double().flatten(0, 1).requires_grad_(True)
        yr = ref(xbr)
        yr.backward(gyb.double().flatten(0, 1))
        cnt = xb.numel() // C
        errs = {"lif gx": e_lif, "lif v": e_v, "bn y": rel(yb.detach().cpu().flatten(0, 1), yr.detach()),
                "bn gw": rel(bn.weight.grad.cpu(), ref.weight.grad), "bn gb": rel(bn.bias.grad.cpu(), ref.bias.grad),
                "bn run_mean": rel(bn.running_mean.cpu(), ref.running_mean)}
        if cnt > 1:   # with a single element per channel the variance is degenerate (0 / 0)
            errs["bn run_var"] = rel(bn.running_var.cpu(), ref.running_var)
        if cnt > 2:   # with two elements the normalised values are +-1 and the input gradient cancels to O(eps): its
            errs["bn gx"] = rel(xbg.grad.cpu().flatten(0, 1), xbr.grad)   # relative error measures fp32 cancellation only
        lim = {"bn gx": 2e-4}
        ok = same and all(v <= lim.get(k, 2e-5) for k, v in errs.items())
        bad += not ok
        print(f"case {case}: T={T} shape={shape} tau={tau} v_th={v_th} v_reset={v_reset} decay={decay} detach={detach}: "
              f"spikes identical {same}, " + " ".join(f"{k} {v:.1e}" for k, v in errs.items()), "" if ok else "<-- CHECK", flush=True)
    print("suspicious cases:", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
