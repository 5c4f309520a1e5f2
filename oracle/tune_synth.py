"""Offline tuning aid for ``spiking-diffusion_b200/synth.py`` -- TEST INFRASTRUCTURE.

Prints per-layer eval-mode firing rates, the number of distinct codes used and logit statistics for the
synthetic parameters, using the CPU oracle.  The constants in synth.py (_ENC/_DEC/_DEN/_GEN_BETA) were
chosen with this script so that rates land in roughly 5-15 %.  Usage: python oracle/tune_synth.py [T]
"""
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import snn_oracle as O  # noqa: E402


def _load_synth():
    spec = importlib.util.spec_from_file_location(
        "sd_synth", os.path.join(os.path.dirname(HERE), "spiking-diffusion_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    S = _load_synth()
    torch.manual_seed(0)
    for K in (128, 512):
        p = S.synth_vqvae_state(0, num_embeddings=K, T=T)
        img = S.synth_images(0, 32)
        tr = O.Trace()
        e, rec, idx = O.vqvae_forward_eval(img.unsqueeze(0).repeat(T, 1, 1, 1, 1), p, trace=tr)
        print(f"[vqvae K={K} T={T}] rates:", {k: round(float(v[0].mean()), 4) for k, v in tr.items() if k != 'feat'})
        print("   distinct codes:", idx.unique().numel(), "of", K, " feat range", float(tr['feat'].min()),
              float(tr['feat'].max()), " recon range", float(rec.min()), float(rec.max()), "recon std", float(rec.std()))
        d = S.synth_denoiser_state(0, num_embeddings=K)
        for frac_masked in (1.0, 0.5, 0.0):
            b = 16
            x = torch.randint(0, K, (b, 1, 7, 7)).float()
            m = torch.rand(b, 1, 7, 7) < frac_masked
            x[m] = K
            t = torch.randint(1, 50, (b,))
            tr = O.Trace()
            lg = O.denoiser_forward(x, t, d, T, trace=tr)
            print(f"[denoiser K={K} masked={frac_masked}] rates:",
                  {k: round(float(v[0].mean()), 4) for k, v in tr.items()},
                  " logits mean/std/min/max: %.3f %.3f %.3f %.3f" % (lg.mean(), lg.std(), lg.min(), lg.max()))
            pr = O.categorical_probs(lg.permute(0, 2, 3, 1))
            print("   max prob mean %.4f  entropy %.3f" % (pr.max(-1).values.mean(),
                                                           -(pr * pr.clamp_min(1e-30).log()).sum(-1).mean()))


if __name__ == "__main__":
    main()
