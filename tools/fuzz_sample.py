"""Random-configuration check of AbsorbingDiffusion.sample against the CPU oracle driven by the same Philox stream
(GPU box): batch, latent size, codebook size, T, temperature, number of steps.  Usage: python tools/fuzz_sample.py [n] [seed]"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import philox, snn_oracle as O  # noqa: E402  (checker)
from spiking_diffusion_b200 import _lib, synth  # noqa: E402
from spiking_diffusion_b200.activation_based import functional  # noqa: E402
from spiking_diffusion_b200.snn_model.vq_diffusion import AbsorbingDiffusion, DummyModel  # noqa: E402


def main():
    import ctypes
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    info = [ctypes.c_int() for _ in range(4)]
    _lib.check(_lib.lib().sd_device_info(*[ctypes.byref(v) for v in info]))
    sms, thr = info[0].value, info[1].value
    tot_imgs = same_imgs = 0
    bad = 0
    for case in range(n_cases):
        T = rng.choice([1, 2, 4, 8])
        hw = rng.choice([4, 5, 7, 8])
        K = rng.choice([32, 48, 100, 128, 200, 512, 1000])
        b = rng.choice([1, 2, 3, 5, 8])
        temp = rng.choice([0.65, 0.9, 1.0])
        steps = rng.choice([hw * hw, hw * hw, max(3, hw * hw // 3)])
        seed = 10 + case
        sd = synth.synth_denoiser_state(case, n_channel=1, num_embeddings=K, num_timesteps=hw * hw)
        den = DummyModel(1, K, T=T)
        functional.set_step_mode(den, "m")
        den.load_state_dict(sd)
        den = den.eval().cuda()
        ab = AbsorbingDiffusion(den, mask_id=K, shape=(hw, hw), n_samples=b)
        x = ab.sample(temp=temp, sample_steps=steps, seed=seed)
        plan = ab.plan(b)
        inc = plan.inc_u + plan.inc_e
        uni = lambda step, n: torch.from_numpy(philox.uniform(seed, step * inc, n, sms, thr))
        expo = lambda step, rows, k: torch.from_numpy(
            philox.exponential(seed, step * inc + plan.inc_u, rows * k, sms, thr)).reshape(rows, k)
        x_ref = O.sample(sd, T, b, (hw, hw), K, K, temp, steps, uni, expo)
        same = int((x.cpu() == x_ref).reshape(b, -1).all(dim=1).sum())
        full = int(x.max()) < K and int(x.min()) >= 0
        tot_imgs += b; same_imgs += same
        ok = full and same >= b - 1
        bad += not ok
        print(f"case {case}: T={T} latent {hw}x{hw} K={K} b={b} temp={temp} steps={steps}: images identical to the oracle's "
              f"{same}/{b}, fully unmasked {full}", "" if ok else "<-- CHECK", flush=True)
    print(f"identical trajectories: {same_imgs}/{tot_imgs} images; suspicious cases: {bad}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
