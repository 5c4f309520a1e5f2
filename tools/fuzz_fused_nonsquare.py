"""Fused conv+BN+LIF layers on NON-SQUARE grids (every other test uses square ones): tcgen05 kernel, CUDA-core spike
kernel and the real-input kernel against the CPU oracle.  Usage: python tools/fuzz_fused_nonsquare.py [n] [seed]"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_conv import make_block, spikes  # noqa: E402
from oracle import snn_oracle as O  # noqa: E402  (checker)
from spiking_diffusion_b200 import _lib, engine  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    for case in range(n_cases):
        T = rng.choice([1, 2, 4, 8])
        H, W = rng.choice([3, 5, 7, 8, 12]), rng.choice([4, 6, 7, 9, 14])
        B = rng.choice([1, 2, 5, 11])
        kind = rng.choice(["tc3x3", "simt3x3", "convT_s2", "conv_s2", "real_const"])
        if kind in ("tc3x3", "simt3x3"):
            cin, cout = rng.choice([16, 64, 80]), rng.choice([16, 48, 128])
            seq, p = make_block(cin, cout, seed=case)
            kw = dict(stride=1, padding=1)
        elif kind == "convT_s2":
            cin, cout = rng.choice([16, 64]), rng.choice([16, 32])
            seq, p = make_block(cin, cout, 3, stride=2, pad=1, transposed=True, op=1, seed=case)
            kw = dict(stride=2, padding=1, transposed=True, output_padding=1)
        elif kind == "conv_s2":
            cin, cout = rng.choice([16, 32]), rng.choice([16, 64])
            seq, p = make_block(cin, cout, 3, stride=2, pad=1, seed=case)
            kw = dict(stride=2, padding=1)
        else:
            cin, cout = rng.choice([1, 2, 3]), rng.choice([16, 64])
            seq, p = make_block(cin, cout, seed=case)
            kw = dict(stride=1, padding=1)
        conv, bn, lif = seq[0], seq[1], seq[2]
        if kind == "real_const":
            x = torch.rand(B, cin, H, W) - 0.3
            x_seq = x.unsqueeze(0).repeat(T, 1, 1, 1, 1)
            lyr = engine.FusedLayer(conv, bn, lif, T=T, B=B, H_in=H, W_in=W, in_kind=_lib.IN_REAL_CONST,
                                    out_kind=_lib.OUT_LIF, impl="simt")
            inp = x.cuda()
        else:
            x_seq = spikes((T, B, cin, H, W), 0.15, 50 + case)
            impl = "tc" if kind == "tc3x3" else "simt"
            lyr = engine.FusedLayer(conv, bn, lif, T=T, B=B, H_in=H, W_in=W, in_kind=_lib.IN_STF, out_kind=_lib.OUT_LIF,
                                    impl=impl)
            inp = engine.stf_from_nchw(x_seq.cuda())
        cur = O.conv_bn(x_seq, p, "c", "b", **kw)
        s_ref, _, h_ref = O.lif_multi_step(cur, return_h=True)
        out = lyr.run(inp, lyr.alloc_out())
        d = lyr.desc
        got = engine.stf_to_nchw(out, T, B, cout, d.H_out, d.W_out).cpu()
        shape_ok = tuple(got.shape) == tuple(s_ref.shape)
        near = torch.cummax(((h_ref - 1.0).abs() <= 1e-4).to(torch.uint8), dim=0).values.bool() if shape_ok else None
        hard = int(((got != s_ref) & ~near).sum()) if shape_ok else -1
        rate = float((got != s_ref).float().mean()) if shape_ok else 1.0
        ok = shape_ok and hard == 0 and rate <= 1e-4
        bad += not ok
        print(f"case {case}: {kind} {cin}->{cout} {H}x{W} T={T} B={B}: shape ok {shape_ok}, flips outside margin {hard}, "
              f"flip rate {rate:.1e}", "" if ok else "<-- CHECK", flush=True)
    print("suspicious cases:", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
