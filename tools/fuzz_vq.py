"""Random-shape check of the vector-quantiser lookup (sd_vq_lookup) against the reference's formula evaluated by torch
in float64 (R/snn_model/vae_model.py:87-95), with the distance-gap margin rule.  Usage: python tools/fuzz_vq.py [n] [seed]"""
import ctypes
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spiking_diffusion_b200 import _lib  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    L = _lib.lib()
    bad = 0
    for case in range(n_cases):
        M = rng.choice([1, 3, 49, 333, 3136, 20000])
        D = rng.choice([1, 4, 16, 16, 24, 64])
        K = rng.choice([1, 2, 7, 64, 128, 512, 1000, 4096])
        g = torch.Generator().manual_seed(case)
        z = (torch.rand(M, D, generator=g) * 2.2).cuda()
        cb = (torch.randn(K, D, generator=g) * 0.7 + 1.0).cuda()
        if K > 2:
            cb[K // 2] = cb[1]   # an exact duplicate: the first index must win
        idx = torch.empty(M, dtype=torch.int64, device="cuda")
        margin = torch.empty(M, dtype=torch.float32, device="cuda")
        rc = L.sd_vq_lookup(_lib.ptr(z), _lib.ptr(cb), _lib.ptr(idx), _lib.ptr(margin), M, D, K, _lib.stream_ptr())
        if rc:
            print(f"case {case}: M={M} D={D} K={K}: rejected: {L.sd_last_error().decode()}")
            continue
        zd, cd = z.double(), cb.double()
        dist = (zd ** 2).sum(1, keepdim=True) + (cd ** 2).sum(1) - 2 * zd @ cd.t()
        ref = dist.argmin(1)
        top2 = dist.topk(min(2, K), dim=1, largest=False).values
        gap = (top2[:, 1] - top2[:, 0]) if K > 1 else torch.full((M,), 1e9, device="cuda", dtype=torch.float64)
        wrong = idx != ref
        # a differing index is acceptable only where the two best distances are within 1e-4 (fp32 evaluation order)
        hard = int((wrong & (gap > 1e-4)).sum())
        dup_ok = bool(((idx != K // 2) | (K <= 2)).all()) if K > 2 else True
        m_ok = K == 1 or float((margin.double() - gap).abs().max()) <= 1e-3
        ok = hard == 0 and dup_ok and m_ok and int(idx.min()) >= 0 and int(idx.max()) < K
        bad += not ok
        print(f"case {case}: M={M} D={D} K={K}: mismatches {int(wrong.sum())} (outside the 1e-4 gap: {hard}), duplicate rule "
              f"{dup_ok}, margin output ok {m_ok}", "" if ok else "<-- CHECK", flush=True)
    print("suspicious cases:", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
