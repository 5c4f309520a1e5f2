"""Seeded synthetic parameters in the reference's ``state_dict`` format.

No trained checkpoint ships with the reference, and a default-initialised network is silent (every LIF
layer fires at 0 %, SURVEY.md finding 5), which would make parity tests vacuous.  This module builds
``state_dict``s with the reference's own keys (R/main.py:199,286 save format; key list in SURVEY.md
section 8(f)) from a seed alone, so that the same parameters can be regenerated bit-identically in this
container, on the GPU box and inside ``oracle/gen_golden.py``:

* conv / conv-transpose weights and biases: PyTorch's default init, U(-1/sqrt(fan_in), 1/sqrt(fan_in)),
  drawn from a seeded CPU ``torch.Generator``;
* BatchNorm running statistics: the analytic per-channel mean/variance of the conv output under an
  i.i.d. model of the layer's input (Bernoulli spikes at an assumed rate, or the real-valued input
  distribution of the three real-input layers), accumulated in float64;
* BatchNorm beta: one constant per layer, tuned offline (``oracle/tune_synth.py``) so that eval-mode firing
  rates land in roughly 5-15 %;
* codebook: sparse non-negative codes matched to the range of the VQ feature
  (1-alpha)*memout(spikes) + alpha*rate.

Only RNG draws and element-wise / float64-sum arithmetic are used (no convolutions), so the result does
not depend on the host's BLAS/oneDNN kernel selection.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

Tensor = torch.Tensor

# Assumed input firing rate and beta per layer (tuned by oracle/tune_synth.py).
_ENC = {"rate": [None, 0.10, 0.10], "beta": [0.2, 0.45, 0.4]}
_DEC = {"rate": [0.10, 0.10], "beta": [0.35, 0.35]}
_GEN_BETA = 0.0
_DEN = {"rate": [None, 0.10, 0.10, 0.10, 0.10], "beta": [0.0, 0.45, 0.45, 0.45, 0.45]}


def _uniform(g: torch.Generator, shape, bound: float) -> Tensor:
    return (torch.rand(shape, generator=g, dtype=torch.float64) * 2.0 - 1.0).mul_(bound).float()


def _conv_init(g, cout, cin, k, transposed=False):
    """nn.Conv2d / nn.ConvTranspose2d default init bounds (kaiming_uniform(a=sqrt(5)) = 1/sqrt(fan_in)).

    For ConvTranspose2d torch computes fan_in from weight.size(1)*k*k with weight [C_in, C_out, k, k],
    i.e. C_out*k*k.
    """
    if transposed:
        fan_in = cout * k * k
        w = _uniform(g, (cin, cout, k, k), 1.0 / math.sqrt(fan_in))
    else:
        fan_in = cin * k * k
        w = _uniform(g, (cout, cin, k, k), 1.0 / math.sqrt(fan_in))
    b = _uniform(g, (cout,), 1.0 / math.sqrt(fan_in))
    return w, b


def _per_out_channel(w: Tensor, transposed: bool):
    """(sum w, sum w^2) per output channel in float64."""
    wd = w.double()
    dims = (0, 2, 3) if transposed else (1, 2, 3)
    return wd.sum(dim=dims), (wd * wd).sum(dim=dims)


def _bn(sd: Dict[str, Tensor], name: str, mean: Tensor, var: Tensor, beta: float):
    c = mean.numel()
    sd[name + ".weight"] = torch.ones(c)
    sd[name + ".bias"] = torch.full((c,), float(beta))
    sd[name + ".running_mean"] = mean.float()
    sd[name + ".running_var"] = var.clamp_min(1e-4).float()
    sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _spike_layer_stats(w, b, rate, transposed=False, tap_frac=1.0):
    s1, s2 = _per_out_channel(w, transposed)
    mean = rate * tap_frac * s1 + b.double()
    var = rate * (1.0 - rate) * tap_frac * s2
    return mean, var


def memout_coef(T: int) -> Tensor:
    """R/snn_model/snn_layers.py:31-34 with n_steps := T."""
    return torch.pow(0.8, torch.arange(T - 1, -1, -1))[:, None, None, None, None]


def synth_vqvae_state(seed: int = 0, in_dim: int = 1, embedding_dim: int = 16, num_embeddings: int = 128,
                      T: int = 4) -> Dict[str, Tensor]:
    """``SNN_VQVAE(in_dim, embedding_dim, num_embeddings, .)`` state dict (R/snn_model/vae_model.py:161-177)."""
    g = torch.Generator().manual_seed(1000 + seed)
    sd: Dict[str, Tensor] = {}
    D = embedding_dim
    # encoder: conv(in,32,k3,s2,p1) BN LIF | conv(32,64,k3,s2,p1) BN LIF | conv(64,D,k1) BN LIF
    spec = [(32, in_dim, 3), (64, 32, 3), (D, 64, 1)]
    for i, (co, ci, k) in enumerate(spec):
        w, b = _conv_init(g, co, ci, k)
        sd[f"encoder.snn_convs.{3 * i}.weight"], sd[f"encoder.snn_convs.{3 * i}.bias"] = w, b
        if i == 0:  # image ~ U(-0.5, 0.5): mean 0, var 1/12  (R/main.py:131)
            s1, s2 = _per_out_channel(w, False)
            mean, var = b.double(), s2 / 12.0
        else:
            mean, var = _spike_layer_stats(w, b, _ENC["rate"][i])
        _bn(sd, f"encoder.snn_convs.{3 * i + 1}", mean, var, _ENC["beta"][i])
    # vector quantiser
    sd["vq_layer.alpha"] = torch.tensor(0.5)
    sd["vq_layer.memout.coef"] = memout_coef(T)
    act = (torch.rand((num_embeddings, D), generator=g, dtype=torch.float64) < 0.15).double()
    mag = torch.rand((num_embeddings, D), generator=g, dtype=torch.float64) * 1.0 + 0.2
    cb = (act * mag).float()
    sd["vq_layer.embeddings.weight"] = cb
    w, b = _conv_init(g, D, D, 1)
    sd["vq_layer.poisson.0.weight"], sd["vq_layer.poisson.0.bias"] = w, b
    mu_z, var_z = cb.double().mean(), cb.double().var(unbiased=False)
    s1, s2 = _per_out_channel(w, False)
    _bn(sd, "vq_layer.poisson.1", mu_z * s1 + b.double(), var_z * s2, _GEN_BETA)
    # decoder: convT(D,64,k3,s2,p1,op1) BN LIF | convT(64,32,...) BN LIF | convT(32,in,k3,s1,p1)
    spec = [(64, D, 3), (32, 64, 3)]
    for i, (co, ci, k) in enumerate(spec):
        w, b = _conv_init(g, co, ci, k, transposed=True)
        sd[f"decoder.snn_convs.{3 * i}.weight"], sd[f"decoder.snn_convs.{3 * i}.bias"] = w, b
        # stride-2 transposed conv: an output pixel sees on average 9/4 of the 9 taps
        mean, var = _spike_layer_stats(w, b, _DEC["rate"][i], transposed=True, tap_frac=0.25)
        _bn(sd, f"decoder.snn_convs.{3 * i + 1}", mean, var, _DEC["beta"][i])
    w, b = _conv_init(g, in_dim, 32, 3, transposed=True)
    sd["decoder.snn_convs.6.weight"], sd["decoder.snn_convs.6.bias"] = w * 1.0, b
    sd["memout.coef"] = memout_coef(T)
    return sd


def synth_denoiser_state(seed: int = 0, n_channel: int = 1, num_embeddings: int = 128,
                         num_timesteps: int = 49) -> Dict[str, Tensor]:
    """``DummyModel(n_channel, num_embeddings)`` state dict (R/snn_model/vq_diffusion.py:158-187)."""
    g = torch.Generator().manual_seed(2000 + seed)
    sd: Dict[str, Tensor] = {}
    K = num_embeddings
    chans = [(64, 2 * n_channel), (128, 64), (256, 128), (512, 256), (256, 512)]
    for i, (co, ci) in enumerate(chans):
        w, b = _conv_init(g, co, ci, 3)
        sd[f"conv{i + 1}.0.weight"], sd[f"conv{i + 1}.0.bias"] = w, b
        if i == 0:
            # input = cat(token id as float, t): token ~ half mask id K, half uniform codes; t ~ U{1..steps}
            wd = w.double()
            mu_x = 0.5 * K + 0.5 * (K - 1) / 2.0
            ex2 = 0.5 * K * K + 0.5 * (K - 1) * (2 * K - 1) / 6.0
            var_x = ex2 - mu_x * mu_x
            mu_t = (num_timesteps + 1) / 2.0
            var_t = (num_timesteps ** 2 - 1) / 12.0
            sx = wd[:, :n_channel].sum(dim=(1, 2, 3))
            st = wd[:, n_channel:].sum(dim=(1, 2, 3))
            sx2 = (wd[:, :n_channel] ** 2).sum(dim=(1, 2, 3))
            mean = mu_x * sx + mu_t * st + b.double()
            var = var_x * sx2 + var_t * st * st   # t is constant over the 3x3 window
        else:
            mean, var = _spike_layer_stats(w, b, _DEN["rate"][i])
        _bn(sd, f"conv{i + 1}.1", mean, var, _DEN["beta"][i])
    w, b = _conv_init(g, K, 256 + 64, 3)
    sd["conv6.0.weight"], sd["conv6.0.bias"] = w * 8.0, b
    return sd


def synth_images(seed: int, b: int, in_dim: int = 1, size: int = 28) -> Tensor:
    """Synthetic 'ToTensor() then -0.5' images (R/load_dataset_snn.py:22-26, R/main.py:131)."""
    g = torch.Generator().manual_seed(3000 + seed)
    return torch.rand((b, in_dim, size, size), generator=g) - 0.5
