"""(a12) margin-conditional sample parity at benchmark scale.

sample() on the GPU against the CPU oracle under the same torch-CUDA Philox stream, at the shapes of BASELINE.json
configs[1..3]: cfg2 (28x28, b=256, T=4, K=128: 32 images of the global stream), cfg3 (CIFAR-shape 8x8 grid, 64 steps,
b=1024: 8 images) and cfg4 (T=8, K=512, a 512-image shard at offset 1024 of a 4096-image global batch: 8 images).
Every image is either identical after every step or its first divergence is attributed to a near-threshold neuron or
a near-tie draw (tests/_sample_parity.py); the reports are written to gpurun_out/r02_sample_parity_<case>.json and
copied to profiles/."""
import pytest
import torch

from _sample_parity import run_case
from spiking_diffusion_b200 import synth
from spiking_diffusion_b200.activation_based import functional
from spiking_diffusion_b200.snn_model.vq_diffusion import AbsorbingDiffusion, DummyModel

pytestmark = pytest.mark.gpu

CASES = {
    "cfg2": dict(T=4, K=128, hw=7, b=256, n_global=256, shard_base=0, check=list(range(0, 256, 8)), temp=1.0, seed=11),
    "cfg3": dict(T=4, K=128, hw=8, b=1024, n_global=1024, shard_base=0, check=[0, 131, 262, 393, 524, 655, 786, 1023],
                 temp=1.0, seed=12),
    "cfg4": dict(T=8, K=512, hw=7, b=512, n_global=4096, shard_base=1024, check=[0, 73, 146, 219, 292, 365, 438, 511],
                 temp=0.65, seed=13),
    # the reference as shipped: T=16, 16 images per call (R/snn_model/vq_diffusion.py:51), T-parallel small-batch mode
    "ref16": dict(T=16, K=128, hw=7, b=16, n_global=16, shard_base=0, check=[0, 5, 10, 15], temp=0.65, seed=14),
}


@pytest.mark.parametrize("name", list(CASES))
def test_sample_matches_oracle_or_diverges_for_a_permitted_reason(name):
    c = dict(CASES[name])
    T, K, hw, b = c["T"], c["K"], c["hw"], c.pop("b")
    dsd = synth.synth_denoiser_state(0, n_channel=1, num_embeddings=K, num_timesteps=hw * hw)
    den = DummyModel(1, K, T=T)
    functional.set_step_mode(den, "m")
    den.load_state_dict(dsd)
    den = den.eval().cuda()
    ab = AbsorbingDiffusion(den, mask_id=K, shape=(hw, hw), n_samples=b)
    rep = run_case(name, den, dsd, ab, **c)
    n = rep["images_checked"]
    print(f"{name}: {rep['identical']} of {n} images identical after all {rep['steps']} steps; "
          f"diverged: {[(d['image'], d['first_step'], d['cause']) for d in rep['diverged']]}")
    assert rep["identical"] + len(rep["diverged"]) == n
    # permitted divergences are rare events: most images must coincide end to end
    assert rep["identical"] >= n - max(1, n // 4), rep
