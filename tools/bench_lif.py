"""HBM roofline of the stand-alone kernels (GPU box): LIFNode.forward, VQ lookup, sampling step."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from spiking_diffusion_b200 import _lib  # noqa: E402
from spiking_diffusion_b200.activation_based import neuron  # noqa: E402


def timeit(fn, reps=20, flush=None):
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts[3:])
    return sum(ts) / len(ts)


def main():
    pk = bench.peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    # (a1) LIF: conv4-sized activation of cfg2 and a large one; 8 B per neuron-timestep + 8 B per neuron
    for name, shape in (("cfg2 den.conv4 output [4,256,512,7,7]", (4, 256, 512, 7, 7)), ("[16,64,512,28,28]", (16, 64, 512, 28, 28))):
        x = torch.rand(shape, device="cuda") * 2
        n = neuron.LIFNode(step_mode="m").eval()
        n(x)
        def run():
            n.reset(); n.v = v0
            return n(x)
        v0 = torch.zeros(shape[1:], device="cuda")
        ms = timeit(run, flush=flush)
        T, N = shape[0], x[0].numel()
        byts = 8 * T * N + 8 * N
        out["lif " + name] = {"ms": round(ms, 4), "GB/s": round(byts / ms / 1e6, 1), "frac_of_measured_hbm": round(byts / ms / 1e6 / pk["hbm"], 3),
                              "algorithmic_bytes": byts}
    # (a12) sampling step: 4*K B per token of logits + 9 B of token state
    L = _lib.lib()
    for K, ntok in ((128, 256 * 49), (512, 4096 * 49)):
        logits = torch.randn(ntok, K, device="cuda")
        xt = torch.full((ntok,), K, dtype=torch.int64, device="cuda")
        um = torch.zeros(ntok, dtype=torch.uint8, device="cuda")
        def run():
            _lib.check(L.sd_sample_step(logits.data_ptr(), xt.data_ptr(), um.data_ptr(), None, ntok, K, 3, 1.0, 1, 0, 4, 0, ntok,
                                        _lib.stream_ptr()))
        ms = timeit(run, flush=flush)
        byts = ntok * (4 * K + 9)
        out[f"sample_step K={K} tokens={ntok}"] = {"ms": round(ms, 4), "GB/s": round(byts / ms / 1e6, 1),
                                                   "frac_of_measured_hbm": round(byts / ms / 1e6 / pk["hbm"], 3)}
    # (a7) VQ lookup
    for K, M in ((128, 64 * 49), (512, 4096 * 49)):
        z = torch.rand(M, 16, device="cuda")
        cb = torch.rand(K, 16, device="cuda")
        idx = torch.empty(M, dtype=torch.int64, device="cuda")
        def run():
            _lib.check(L.sd_vq_lookup(z.data_ptr(), cb.data_ptr(), idx.data_ptr(), None, M, 16, K, _lib.stream_ptr()))
        ms = timeit(run, flush=flush)
        out[f"vq_lookup K={K} M={M}"] = {"ms": round(ms, 4), "GMAC/s": round(M * K * 16 / ms / 1e6, 1)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
