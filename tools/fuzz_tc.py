"""Random-shape cross-check of the tcgen05 conv kernel against the CUDA-core kernel (GPU box): batch sizes around the
tile / pair boundaries, T = 1..16 including non-multiples of 4, channel counts that leave padded N tiles, grids 3..28.
Both weight representations of the tensor-core kernel are exercised: two fp16 terms (kind::f16, fp16 STF spikes) and three
int8 digits (kind::i8, u8 STF8 spikes; needs an even T and C_in % 32 == 0).
Usage: python tools/fuzz_tc.py [n_cases] [seed]"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_conv import make_block, spikes, tc_layer  # noqa: E402
from spiking_diffusion_b200 import _lib, engine  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    worst, bad = 0.0, []
    for case in range(n_cases):
        T = rng.choice([1, 2, 3, 4, 4, 8, 8, 12, 16])
        ns = rng.choice([2, 3, 3])
        H = rng.choice([3, 5, 7, 7, 8, 14, 28])
        cin = rng.choice([16, 32, 48, 64, 128, 320])
        cout = rng.choice([16, 32, 48, 64, 128, 144, 256])
        rows_target = rng.choice([60, 128, 129, 256, 300, 512, 1000, 3000])
        B = max(1, rows_target // (H * H))
        conc = rng.choice([1, 1, 3])
        seq, p = make_block(cin, cout, seed=case)
        kw = dict(concurrent=conc)
        try:
            a = tc_layer(seq, T, B, H, _lib.OUT_LIF, ns, **kw)
        except ValueError as e:   # unsupported by the tc kernel: fine, but say so
            print(f"case {case}: nsplit={ns} T={T} H={H} cin={cin} cout={cout} B={B}: tc unsupported ({e})")
            continue
        b = tc_layer(seq, T, B, H, _lib.OUT_LIF, 2, impl="simt")
        s_in = spikes((T, B, cin, H, H), rng.choice([0.05, 0.15, 0.4]), 100 + case)
        x = engine.stf_from_nchw(s_in.cuda())
        xa = engine.stf8_from_nchw(s_in.cuda()) if ns == 3 else x
        va, vb = a.alloc_state(), b.alloc_state()
        oa, ob, sa, sb = a.alloc_out(), b.alloc_out(), a.alloc_sum(), b.alloc_sum()
        for _ in range(2):   # second call continues from the carried state
            a.run(xa, oa, out_sum=sa, v=va); b.run(x, ob, out_sum=sb, v=vb)
        ga = engine.stf8_to_nchw(oa, T, B, cout, H, H) if ns == 3 else engine.stf_to_nchw(oa, T, B, cout, H, H)
        gb = engine.stf_to_nchw(ob, T, B, cout, H, H)
        flips = float((ga != gb).float().mean())
        cnt_ok = torch.equal(engine.stf_to_nchw(sa, 1, B, cout, H, H)[0], ga.sum(0))
        key = _lib.lib().sd_conv_weight_layout_tc(__import__("ctypes").byref(a.desc))
        tag = f"case {case}: nsplit={ns} T={T} H={H} cin={cin} cout={cout} B={B} conc={conc} layout N={key >> 16} K={(key >> 4) & 0xfff} pair={key & 1}"
        worst = max(worst, flips)
        ok = flips <= 2e-4 and cnt_ok
        print(tag, f"flip rate {flips:.1e}", "" if cnt_ok else "T-SUM MISMATCH", "" if ok else "<-- CHECK", flush=True)
        if not ok:
            bad.append(tag)
    print(f"worst flip rate {worst:.2e}; suspicious cases: {len(bad)}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
