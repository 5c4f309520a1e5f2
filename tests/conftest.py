"""Shared test plumbing.  `-m "not gpu"` tests run in the CPU build container; `-m gpu` tests are the parity tests
proper and call the CUDA path through the C-ABI on a real B200."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

# Tolerances of BASELINE.json:north_star / BASELINE.md section 5
SPIKE_MARGIN = 1e-4      # spikes/indices must match bit-exactly where the reference margin exceeds this
FLIP_RATE_MAX = 1e-4     # overall spike-flip rate
IMAGE_TOL = 1e-3         # decoded images, max-abs


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device (GPU tests run under gpurun)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The product path fails loudly without its CUDA extension; build it once per session."""
    import spiking_diffusion_b200 as sd
    if not os.path.exists(sd._lib.LIB_PATH):
        sd.build()
    return sd


def unpack(bits: np.ndarray, shape) -> torch.Tensor:
    n = int(np.prod(shape))
    return torch.from_numpy(np.unpackbits(bits)[:n].reshape(tuple(int(s) for s in shape)).astype(np.float32))


def golden(name: str):
    return np.load(os.path.join(GOLD, name))


def flip_stats(got: torch.Tensor, ref: torch.Tensor, near: torch.Tensor = None):
    """(flip rate, number of flips outside the near-threshold set)."""
    diff = got != ref
    rate = float(diff.float().mean())
    hard = int((diff & ~near.bool()).sum()) if near is not None else int(diff.sum())
    return rate, hard


# ---- model builders shared by the GPU tests ------------------------------------------------------------
def make_vqvae(T, K=128, seed=0, device="cuda"):
    from spiking_diffusion_b200 import synth
    from spiking_diffusion_b200.activation_based import functional
    from spiking_diffusion_b200.snn_model.vae_model import SNN_VQVAE
    sd = synth.synth_vqvae_state(seed, num_embeddings=K, T=T)
    m = SNN_VQVAE(1, 16, K, torch.tensor(1.0), T=T)
    functional.set_step_mode(m, "m")
    m.load_state_dict(sd)
    return m.eval().to(device), sd


def make_denoiser(T, K=128, seed=0, device="cuda"):
    from spiking_diffusion_b200 import synth
    from spiking_diffusion_b200.activation_based import functional
    from spiking_diffusion_b200.snn_model.vq_diffusion import DummyModel
    sd = synth.synth_denoiser_state(seed, num_embeddings=K)
    m = DummyModel(1, K, T=T)
    functional.set_step_mode(m, "m")
    m.load_state_dict(sd)
    return m.eval().to(device), sd


def assert_spikes_match(got: torch.Tensor, ref: torch.Tensor, h_ref: torch.Tensor, name: str, v_th: float = 1.0):
    """north_star rule: bit-exact wherever |h - v_th| > 1e-4; overall flip rate <= 1e-4.

    A neuron whose potential came within the margin at timestep t may legitimately differ at every LATER timestep
    too (the flipped spike resets its membrane), so the near-threshold set is accumulated along time."""
    got, ref = got.cpu(), ref.cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    near = torch.cummax(((h_ref.cpu() - v_th).abs() <= SPIKE_MARGIN).to(torch.uint8), dim=0).values.bool()
    diff = got != ref
    hard = int((diff & ~near).sum())
    rate = float(diff.float().mean())
    assert hard == 0, f"{name}: {hard} spikes differ outside the 1e-4 margin (flip rate {rate:.2e})"
    assert rate <= FLIP_RATE_MAX, f"{name}: flip rate {rate:.2e} > {FLIP_RATE_MAX}"
    return rate
