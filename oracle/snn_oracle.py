"""CPU oracle for the Spiking-Diffusion hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU fp32 restatement of the reference's algorithm for the path named by
BASELINE.json:north_star.  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``spiking-diffusion_b200/`` imports it, and the product path never falls back to it.

Citation prefixes:
  R/  = /root/reference/Spiking-Diffusion-release/
  SJ/ = a file inside R/spikingjelly.zip (the vendored SpikingJelly snapshot)

The reference hard-codes T=16, the 7x7 latent grid, 49 diffusion steps and n_samples=16
(SURVEY.md finding 1-2).  Every function here takes those as arguments; the substitution sites are
listed next to each function.  ``tests/test_oracle_pin.py`` proves the restatement bit-identical to the
reference imported as-is (T=16, 7x7) whenever /root/reference is present, and
``tests/test_oracle_golden.py`` checks it against the committed outputs of that reference
(``tests/golden/*.npz``, produced by ``oracle/gen_golden.py``).  Pin status: PINNED for the VQ-VAE
forward, the denoiser forward and the LIF/memout/PSP/tie-break known-answer vectors; the CUDA Philox
stream used by ``sample`` is pinned on the GPU box against torch's own CUDA generator
(``tests/test_gpu_sampling.py``) because this container has no GPU to produce fixtures from.

Weights are passed as a flat ``dict[str, Tensor]`` that uses the reference's own ``state_dict`` keys.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, used by SJ/activation_based/layer.py:423-465


# --------------------------------------------------------------------------------------------------
# LIF neuron
# --------------------------------------------------------------------------------------------------
def lif_multi_step(x_seq: Tensor, v: Optional[Tensor] = None, tau: float = 2.0, v_threshold: float = 1.0,
                   v_reset: Optional[float] = 0.0, decay_input: bool = True,
                   return_h: bool = False):
    """Multi-step LIF forward, eval branch.

    Follows SJ/activation_based/neuron.py:799-809 (hard reset, decay_input), :778-787 (soft reset),
    the no-decay-input variants next to them, and the state initialisation of
    ``BaseNode.v_float_to_tensor`` (neuron.py:260-263): ``v`` starts as ``full_like(x[0], v_reset)``
    (0. for soft reset).  Returns ``(spike_seq, v_final)`` and, if asked, the pre-fire potential
    ``h_seq`` used by the margin-conditional parity rule.
    """
    if v is None:
        v = torch.full_like(x_seq[0], 0.0 if v_reset is None else float(v_reset))
    spike_seq = torch.zeros_like(x_seq)
    h_seq = torch.empty_like(x_seq) if return_h else None
    for t in range(x_seq.shape[0]):
        if v_reset is None:
            if decay_input:
                v = v + (x_seq[t] - v) / tau
            else:
                v = v * (1.0 - 1.0 / tau) + x_seq[t]
        else:
            if decay_input:
                v = v + (x_seq[t] - (v - v_reset)) / tau
            else:
                v = v - (v - v_reset) / tau + x_seq[t]
        if return_h:
            h_seq[t] = v
        spike = (v >= v_threshold).to(x_seq)
        if v_reset is None:
            v = v - spike * v_threshold
        else:
            v = v_reset * spike + (1.0 - spike) * v
        spike_seq[t] = spike
    if return_h:
        return spike_seq, v, h_seq
    return spike_seq, v


class _ATanSpike(torch.autograd.Function):
    """Heaviside forward, ATan surrogate backward: SJ/activation_based/surrogate.py:663-678."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.save_for_backward(x)
        ctx.alpha = alpha
        return (x >= 0).to(x)

    @staticmethod
    def backward(ctx, grad_output):
        x, = ctx.saved_tensors
        return ctx.alpha / 2 / (1 + (math.pi / 2 * ctx.alpha * x).pow(2)) * grad_output, None


def lif_multi_step_train(x_seq: Tensor, v: Optional[Tensor] = None, tau: float = 2.0, v_threshold: float = 1.0,
                         v_reset: Optional[float] = 0.0, decay_input: bool = True, detach_reset: bool = False,
                         alpha: float = 2.0):
    """Differentiable training branch: BaseNode.multi_step_forward -> single_step_forward
    (SJ/activation_based/neuron.py:244-258,210-242), neuronal_charge (:726-743), neuronal_fire (:161-177),
    neuronal_reset (:179-205).  Returns (spike_seq, v_last); gradients flow through the ATan surrogate."""
    if v is None:
        v = torch.full_like(x_seq[0], 0.0 if v_reset is None else float(v_reset))
    out = []
    for t in range(x_seq.shape[0]):
        x = x_seq[t]
        if decay_input:
            v = v + (x - v) / tau if (v_reset is None or v_reset == 0.0) else v + (x - (v - v_reset)) / tau
        else:
            v = v * (1.0 - 1.0 / tau) + x if (v_reset is None or v_reset == 0.0) else v - (v - v_reset) / tau + x
        spike = _ATanSpike.apply(v - v_threshold, alpha)
        spike_d = spike.detach() if detach_reset else spike
        if v_reset is None:
            v = v - spike_d * v_threshold
        else:
            v = (1.0 - spike_d) * v + spike_d * v_reset
        out.append(spike)
    return torch.stack(out), v


# --------------------------------------------------------------------------------------------------
# Read-out layers
# --------------------------------------------------------------------------------------------------
def memout_coef(T: int) -> Tensor:
    """R/snn_model/snn_layers.py:31-34 with ``n_steps := T``: coef[t] = 0.8 ** (T-1-t), fp32."""
    arr = torch.arange(T - 1, -1, -1)
    return torch.pow(0.8, arr)[:, None, None, None, None]


def memout(x: Tensor) -> Tensor:
    """R/snn_model/snn_layers.py:36-41: sum_t coef[t] * x[t] over the leading time axis."""
    return torch.sum(x * memout_coef(x.shape[0]).to(x), dim=0)


def psp(inputs: Tensor, tau_s: float = 2.0) -> Tensor:
    """R/snn_model/snn_layers.py:6-26: first-order low-pass of a spike train."""
    syn = torch.zeros_like(inputs[0])
    out = []
    for t in range(inputs.shape[0]):
        syn = syn + (inputs[t] - syn) / tau_s
        out.append(syn)
    return torch.stack(out)


# --------------------------------------------------------------------------------------------------
# conv / BN in multi-step mode
# --------------------------------------------------------------------------------------------------
def _seq(fn: Callable[[Tensor], Tensor], x_seq: Tensor) -> Tensor:
    """SJ/activation_based/functional.py:680-688: flatten [T,N] -> [T*N], apply, un-flatten."""
    T, N = x_seq.shape[0], x_seq.shape[1]
    y = fn(x_seq.flatten(0, 1))
    return y.view(T, N, *y.shape[1:])


def conv_bn(x_seq: Tensor, p: Params, conv: str, bn: Optional[str], stride: int = 1, padding: int = 0,
            transposed: bool = False, output_padding: int = 0) -> Tensor:
    """layer.Conv2d / layer.ConvTranspose2d followed by layer.BatchNorm2d in 'm' mode, eval.

    SJ/activation_based/layer.py:164-173 (conv), :316-325 (transposed conv), :458-465 (BN with running
    statistics, eps 1e-5).
    """
    w, b = p[conv + ".weight"], p.get(conv + ".bias")
    if transposed:
        y = _seq(lambda z: F.conv_transpose2d(z, w, b, stride=stride, padding=padding,
                                             output_padding=output_padding), x_seq)
    else:
        y = _seq(lambda z: F.conv2d(z, w, b, stride=stride, padding=padding), x_seq)
    if bn is not None:
        y = _seq(lambda z: F.batch_norm(z, p[bn + ".running_mean"], p[bn + ".running_var"],
                                        p[bn + ".weight"], p[bn + ".bias"], False, 0.1, BN_EPS), y)
    return y


class Trace(dict):
    """Optional per-layer record: ``trace[name] = (spikes, h_seq)`` for the margin-conditional rule."""


def _layer(x_seq, p, conv, bn, trace, name, **kw):
    cur = conv_bn(x_seq, p, conv, bn, **kw)
    if trace is not None:
        s, _, h = lif_multi_step(cur, return_h=True)
        trace[name] = (s, h)
        return s
    return lif_multi_step(cur)[0]


# --------------------------------------------------------------------------------------------------
# VQ-SVAE
# --------------------------------------------------------------------------------------------------
def encoder_forward(x_seq: Tensor, p: Params, prefix: str = "encoder.", trace: Optional[Trace] = None) -> Tensor:
    """R/snn_model/vae_model.py:101-129: conv(k3,s2,p1)-BN-LIF, conv(k3,s2,p1)-BN-LIF, conv(k1)-BN-LIF."""
    q = prefix + "snn_convs."
    x = _layer(x_seq, p, q + "0", q + "1", trace, "enc1", stride=2, padding=1)
    x = _layer(x, p, q + "3", q + "4", trace, "enc2", stride=2, padding=1)
    x = _layer(x, p, q + "6", q + "7", trace, "enc3", stride=1, padding=0)
    return x


def vq_feature(x_seq: Tensor, alpha: Tensor) -> Tensor:
    """R/snn_model/vae_model.py:42-46 with ``num_step := T``: (1-a)*memout(x) + a*sum_t x/T, NHWC-flat."""
    T = x_seq.shape[0]
    x_memout = (1 - alpha) * memout(x_seq) + alpha * torch.sum(x_seq, dim=0) / T
    x_memout = x_memout.permute(0, 2, 3, 1).contiguous()
    return x_memout


def vq_distances(flat_x: Tensor, codebook: Tensor) -> Tensor:
    """R/snn_model/vae_model.py:89-93: (|z|^2 + |e|^2) - 2 z.e^T, in that order, fp32."""
    return (torch.sum(flat_x ** 2, dim=1, keepdim=True) + torch.sum(codebook ** 2, dim=1)
            - 2.0 * torch.matmul(flat_x, codebook.t()))


def vq_code_indices(flat_x: Tensor, codebook: Tensor) -> Tensor:
    """R/snn_model/vae_model.py:87-95: argmin over codes (first index on exact ties)."""
    return torch.argmin(vq_distances(flat_x, codebook), dim=1)


def vq_poisson(quantized_nchw: Tensor, p: Params, T: int, prefix: str = "vq_layer.",
               trace: Optional[Trace] = None) -> Tensor:
    """R/snn_model/vae_model.py:54-57 (and main.py:392-395): repeat(T) -> conv1x1 -> BN -> LIF."""
    q = quantized_nchw.unsqueeze(0).repeat(T, 1, 1, 1, 1)
    return _layer(q, p, prefix + "poisson.0", prefix + "poisson.1", trace, "gen", stride=1, padding=0)


def vq_forward_eval(x_seq: Tensor, p: Params, prefix: str = "vq_layer.", trace: Optional[Trace] = None):
    """R/snn_model/vae_model.py:40-58, eval branch.  Returns (spikes[T,N,D,h,w], indices[N*h*w], feature)."""
    T = x_seq.shape[0]
    cb = p[prefix + "embeddings.weight"]
    feat = vq_feature(x_seq, p[prefix + "alpha"])
    flat = feat.reshape(-1, cb.shape[1])
    idx = vq_code_indices(flat, cb)
    quant = F.embedding(idx, cb).view_as(feat).permute(0, 3, 1, 2).contiguous()
    spikes = vq_poisson(quant, p, T, prefix, trace)
    return spikes, idx, feat


def decoder_forward(e_seq: Tensor, p: Params, prefix: str = "decoder.", trace: Optional[Trace] = None) -> Tensor:
    """R/snn_model/vae_model.py:131-159: convT(k3,s2,p1,op1)-BN-LIF x2, convT(k3,s1,p1) (no BN/LIF)."""
    q = prefix + "snn_convs."
    x = _layer(e_seq, p, q + "0", q + "1", trace, "dec1", stride=2, padding=1, transposed=True, output_padding=1)
    x = _layer(x, p, q + "3", q + "4", trace, "dec2", stride=2, padding=1, transposed=True, output_padding=1)
    return conv_bn(x, p, q + "6", None, stride=1, padding=1, transposed=True, output_padding=0)


def vqvae_forward_eval(x_seq: Tensor, p: Params, trace: Optional[Trace] = None):
    """R/snn_model/vae_model.py:179-187, eval branch: returns (e, x_recon, indices)."""
    z = encoder_forward(x_seq, p, trace=trace)
    e, idx, feat = vq_forward_eval(z, p, trace=trace)
    if trace is not None:
        trace["feat"] = feat
    x_recon = torch.tanh(memout(decoder_forward(e, p, trace=trace)))
    return e, x_recon, idx


def decode_indices(idx_bhw: Tensor, p: Params, T: int, trace: Optional[Trace] = None) -> Tensor:
    """R/main.py:388-399: quantize -> NCHW -> repeat(T) -> poisson -> decoder -> tanh(memout)."""
    z = F.embedding(idx_bhw, p["vq_layer.embeddings.weight"]).permute(0, 3, 1, 2).contiguous()
    e = vq_poisson(z, p, T, trace=trace)
    return torch.tanh(memout(decoder_forward(e, p, trace=trace)))


def to_uint8(pred: Tensor) -> Tensor:
    """R/main.py:401: clip(pred + 0.5, 0, 1) * 255 -> uint8 (truncating cast, as numpy's astype)."""
    return (torch.clamp(pred + 0.5, 0.0, 1.0) * 255).to(torch.uint8)


# --------------------------------------------------------------------------------------------------
# Denoiser (DummyModel) and the absorbing-state sampler
# --------------------------------------------------------------------------------------------------
def denoiser_forward(x: Tensor, t: Tensor, p: Params, T: int, trace: Optional[Trace] = None) -> Tensor:
    """R/snn_model/vq_diffusion.py:189-208 with ``16 := T``.

    ``x`` is [b,1,h,w] float token ids (mask id = K), ``t`` is [b] long.  Returns logits [b,K,h,w].
    """
    tt = torch.ones_like(x) * (t.unsqueeze(1).unsqueeze(2).unsqueeze(3))
    xin = torch.cat((x, tt), dim=1).unsqueeze(0).repeat(T, 1, 1, 1, 1)
    x1 = _layer(xin, p, "conv1.0", "conv1.1", trace, "den1", stride=1, padding=1)
    x2 = _layer(x1, p, "conv2.0", "conv2.1", trace, "den2", stride=1, padding=1)
    x3 = _layer(x2, p, "conv3.0", "conv3.1", trace, "den3", stride=1, padding=1)
    x4 = _layer(x3, p, "conv4.0", "conv4.1", trace, "den4", stride=1, padding=1)
    x5 = _layer(x4, p, "conv5.0", "conv5.1", trace, "den5", stride=1, padding=1)
    x6 = conv_bn(torch.cat((x5, x1), dim=2), p, "conv6.0", None, stride=1, padding=1)
    return torch.sum(x6, dim=0) / T


def categorical_probs(logits: Tensor) -> Tensor:
    """torch.distributions.Categorical(logits=...).probs: normalise by logsumexp, then softmax.

    TORCH/distributions/categorical.py:78 and utils.py:89-100 (``logits_to_probs``).
    """
    norm = logits - logits.logsumexp(dim=-1, keepdim=True)
    return F.softmax(norm, dim=-1)


def sample(p: Params, T: int, b: int, hw: Tuple[int, int], K: int, mask_id: int, temp: float,
           sample_steps: int, uniform_fn: Callable[[int, int], Tensor],
           exponential_fn: Callable[[int, int, int], Tensor],
           denoise: Optional[Callable[[Tensor, Tensor], Tensor]] = None,
           record: Optional[list] = None) -> Tensor:
    """R/snn_model/vq_diffusion.py:103-142 with ``n_samples := b``, ``7,7 := h,w`` and device-agnostic.

    The two random draws per step are injected so the same routine serves torch's CUDA Philox stream
    (``oracle/philox.py``) or any other stream:
      ``uniform_fn(step, numel)``          -> fp32 [numel]   (``torch.rand_like(x_t.float())``, :118)
      ``exponential_fn(step, rows, K)``    -> fp32 [rows,K]  (``exponential_`` inside ``multinomial``,
                                                             TORCH/distributions/categorical.py:147-148)
    ``Categorical.sample`` with one draw is argmax(probs / q), q ~ Exp(1).
    The denoiser state is reset every step (:129), so each call starts from v = 0.
    """
    h, w = hw
    x_t = torch.ones(b, 1, h, w).long() * mask_id
    unmasked = torch.zeros_like(x_t).bool()
    if denoise is None:
        denoise = lambda xf, tt: denoiser_forward(xf, tt, p, T)
    for step, t in enumerate(reversed(range(1, sample_steps + 1))):
        tvec = torch.full((b,), t, dtype=torch.long)
        t_mask = tvec.reshape(b, 1, 1, 1).expand(b, 1, h, w)
        u = uniform_fn(step, b * h * w).reshape(b, 1, h, w)
        changes = u < 1 / t_mask.float()
        changes = torch.bitwise_xor(changes, torch.bitwise_and(changes, unmasked))
        unmasked = torch.bitwise_or(unmasked, changes)
        logits = denoise(x_t.float(), tvec).permute(0, 2, 3, 1)
        logits = logits / temp
        probs = categorical_probs(logits).reshape(-1, K)
        q = exponential_fn(step, b * h * w, K)
        x0_hat = torch.argmax(probs / q, dim=-1).reshape(b, h, w).unsqueeze(1)
        if record is not None:
            record.append({"t": t, "changes": changes.clone(), "logits": logits.clone(), "x0_hat": x0_hat.clone(),
                           "probs": probs, "q": q})
        x_t[changes] = x0_hat[changes]
        if record is not None:
            record[-1]["x_t"] = x_t.clone()     # token grid after this step (trajectory comparison)
    return x_t


# --------------------------------------------------------------------------------------------------
# BN folding used by the fused CUDA layers (so tests can state the expected pre-activation)
# --------------------------------------------------------------------------------------------------
def bn_affine(p: Params, conv: str, bn: Optional[str]) -> Tuple[Tensor, Tensor]:
    """Per-channel (scale, shift) such that BN(conv_nobias(x) + b) = conv_nobias(x)*scale + shift.

    The reference's own fold helper asserts ``conv.bias is None`` (SJ/activation_based/functional.py:731)
    while every reference conv carries a bias, so the fold is re-derived:
      scale = gamma / sqrt(var + eps), shift = (b - mean) * scale + beta.
    """
    b = p[conv + ".bias"]
    if bn is None:
        n = b.numel()
        return torch.ones(n), b.clone()
    inv = torch.rsqrt(p[bn + ".running_var"].double() + BN_EPS)
    scale = p[bn + ".weight"].double() * inv
    shift = (b.double() - p[bn + ".running_mean"].double()) * scale + p[bn + ".bias"].double()
    return scale.float(), shift.float()


def spike_margin(h_seq: Tensor, v_threshold: float = 1.0) -> Tensor:
    """|h - v_th|: a spike may legitimately differ only where this is below the stated tolerance."""
    return (h_seq - v_threshold).abs()


def vq_margin(flat_x: Tensor, codebook: Tensor) -> Tensor:
    """Gap between the best and second-best code distance (the VQ analogue of ``spike_margin``)."""
    d = vq_distances(flat_x, codebook)
    top2 = torch.topk(d, 2, dim=1, largest=False).values
    return top2[:, 1] - top2[:, 0]


# --------------------------------------------------------------------------------------------------
# Training branches: differentiable restatements, gradients by torch autograd
# --------------------------------------------------------------------------------------------------
def conv_bn_train(x_seq: Tensor, p: Params, conv: str, bn: Optional[str], stride: int = 1, padding: int = 0,
                  transposed: bool = False, output_padding: int = 0) -> Tensor:
    """Same as conv_bn but BatchNorm in training mode: batch statistics over T*N*H*W (layer.py:458-465 ->
    F.batch_norm(training=True)); the running buffers in ``p`` are updated in place like nn.BatchNorm2d does."""
    w, b = p[conv + ".weight"], p.get(conv + ".bias")
    if transposed:
        y = _seq(lambda z: F.conv_transpose2d(z, w, b, stride=stride, padding=padding,
                                             output_padding=output_padding), x_seq)
    else:
        y = _seq(lambda z: F.conv2d(z, w, b, stride=stride, padding=padding), x_seq)
    if bn is not None:
        y = _seq(lambda z: F.batch_norm(z, p[bn + ".running_mean"], p[bn + ".running_var"], p[bn + ".weight"],
                                        p[bn + ".bias"], True, 0.1, BN_EPS), y)
    return y


def _layer_train(x_seq, p, conv, bn, **kw):
    return lif_multi_step_train(conv_bn_train(x_seq, p, conv, bn, **kw))[0]


def vqvae_forward_train(x_seq: Tensor, image: Tensor, p: Params, data_variance, commitment_cost: float = 0.25):
    """R/snn_model/vae_model.py:179-196 training branch with VectorQuantizer.forward :40-85 (num_step := T).
    Returns (e_q_loss, recon_loss, real_recon_loss)."""
    T = x_seq.shape[0]
    q = "encoder.snn_convs."
    z = _layer_train(x_seq, p, q + "0", q + "1", stride=2, padding=1)
    z = _layer_train(z, p, q + "3", q + "4", stride=2, padding=1)
    z = _layer_train(z, p, q + "6", q + "7")
    alpha, cb = p["vq_layer.alpha"], p["vq_layer.embeddings.weight"]
    x_memout = (1 - alpha) * memout(z) + alpha * torch.sum(z, dim=0) / T
    x_memout = x_memout.permute(0, 2, 3, 1).contiguous()
    flat = x_memout.reshape(-1, cb.shape[1])
    idx = vq_code_indices(flat, cb)
    quantized = F.embedding(idx, cb).view_as(x_memout)
    q_latent_loss = F.mse_loss(quantized, x_memout.detach())
    e_latent_loss = F.mse_loss(x_memout, quantized.detach())
    loss_1 = q_latent_loss + commitment_cost * e_latent_loss
    quantized = x_memout + (quantized - x_memout).detach()
    quantized = quantized.permute(0, 3, 1, 2).contiguous().unsqueeze(0).repeat(T, 1, 1, 1, 1)
    quantized = _layer_train(quantized, p, "vq_layer.poisson.0", "vq_layer.poisson.1")
    q2 = torch.mean((psp(quantized) - psp(z.detach())) ** 2)
    e2 = torch.mean((psp(quantized.detach()) - psp(z)) ** 2)
    e_q_loss = loss_1 + q2 + commitment_cost * e2
    q = "decoder.snn_convs."
    x = _layer_train(quantized, p, q + "0", q + "1", stride=2, padding=1, transposed=True, output_padding=1)
    x = _layer_train(x, p, q + "3", q + "4", stride=2, padding=1, transposed=True, output_padding=1)
    x = conv_bn_train(x, p, q + "6", None, stride=1, padding=1, transposed=True)
    x_recon = torch.tanh(memout(x))
    real = F.mse_loss(x_recon, image)
    return e_q_loss, real / data_variance, real


def denoiser_forward_train(x: Tensor, t: Tensor, p: Params, T: int) -> Tensor:
    """R/snn_model/vq_diffusion.py:189-208 in training mode (batch-statistics BN, surrogate-gradient LIF)."""
    tt = torch.ones_like(x) * (t.unsqueeze(1).unsqueeze(2).unsqueeze(3))
    xin = torch.cat((x, tt), dim=1).unsqueeze(0).repeat(T, 1, 1, 1, 1)
    x1 = _layer_train(xin, p, "conv1.0", "conv1.1", stride=1, padding=1)
    x5 = x1
    for i in (2, 3, 4, 5):
        x5 = _layer_train(x5, p, f"conv{i}.0", f"conv{i}.1", stride=1, padding=1)
    x6 = conv_bn_train(torch.cat((x5, x1), dim=2), p, "conv6.0", None, stride=1, padding=1)
    return torch.sum(x6, dim=0) / T


def diffusion_train_loss(logits: Tensor, x_0_ignore: Tensor, t: Tensor, num_timesteps: int) -> Tensor:
    """Re-weighted ELBO of R/snn_model/vq_diffusion.py:86-101 given the denoiser logits."""
    b, K = logits.shape[0], logits.shape[1]
    n_tok = logits.shape[2] * logits.shape[3]
    ce = F.cross_entropy(logits.reshape(b, K, n_tok), x_0_ignore.reshape(b, n_tok).long(), ignore_index=-1,
                         reduction="none").sum(1)
    weight = 1 - (t / num_timesteps)
    return (weight * ce / (math.log(2) * n_tok)).mean()
