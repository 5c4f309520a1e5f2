"""VQ-SVAE modules with the reference's constructor signatures, attribute names and ``state_dict`` keys
(mirrors R/snn_model/vae_model.py:22-196), executing on the fused sm_100a kernels.

Differences from the reference, all additive:
* ``T`` (``num_step``) is a keyword argument, default 16 -- the reference hard-codes 16 (vae_model.py:29,42,56).
* ``VectorQuantizer.forward_with_loss`` offers the (quantized, loss, indices) 3-tuple named in the north star;
  ``forward`` keeps the reference's mode-dependent 2-tuple because its callers unpack two values
  (vae_model.py:184,189).
* training mode runs layer by layer through the autograd-capable kernels
  (conv forward / adjoint / weight-gradient, train-mode BatchNorm, surrogate-gradient LIF); the loss algebra
  (MSE, straight-through estimator, PSP, tanh) is plain tensor arithmetic as in the reference.
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib, engine
from .._lib import check, lib, on_device_of, ptr, stream_ptr
from ..activation_based import base, layer, neuron, surrogate
from .snn_layers import PSP, MembraneOutputLayer


def _no_training(mod):
    if mod.training:
        raise NotImplementedError("training mode is not implemented in this round (SURVEY.md section 8(f) rank 1); "
                                  "call .eval() -- the sampling / reconstruction path is eval-only in the reference too")


class VectorQuantizer(nn.Module):
    def __init__(self, embedding_dim, num_embeddings, commitment_cost, T: int = 16):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.num_embeddings = num_embeddings
        self.commitment_cost = commitment_cost
        self.memout = MembraneOutputLayer(T)
        self.num_step = T
        self.psp = PSP()
        self.alpha = nn.Parameter(torch.tensor(0.5))
        self.embeddings = nn.Embedding(self.num_embeddings, self.embedding_dim)
        self.poisson = layer.SpikingSequential(
            layer.Conv2d(in_channels=embedding_dim, out_channels=embedding_dim, kernel_size=1),
            layer.BatchNorm2d(embedding_dim),
            neuron.LIFNode(surrogate_function=surrogate.ATan()),
        )

    @on_device_of
    def feature(self, x: torch.Tensor) -> torch.Tensor:
        """(1-alpha)*memout(x) + alpha*sum_t x/T as fp32 [N, h, w, D]   (vae_model.py:42-44)."""
        T, N, D, h, w = x.shape
        stf = engine.stf_from_nchw(x)
        z = torch.empty((N, h, w, D), dtype=torch.float32, device=x.device)
        check(lib().sd_vq_feature(ptr(stf), ptr(self.alpha.detach()), ctypes.cast(self.memout.coef_host(T), ctypes.c_void_p),
                                  ptr(z), T, N, D, h, w, stream_ptr()))
        return z

    @on_device_of
    def forward(self, x: torch.Tensor):
        """x: [T, N, D, h, w] spikes.  eval -> (spikes [T,N,D,h,w], indices [N*h*w] int64)  (vae_model.py:53-58);
        train -> (spikes, loss_1 + loss_2)  (vae_model.py:61-85)."""
        if x.dim() != 5:
            raise ValueError(f"expected x with shape [T, N, C, H, W], but got x with shape {x.shape}!")
        T = x.shape[0]
        if self.training:
            return self._forward_train(x)
        x_memout = self.feature(x)
        flat_x = x_memout.reshape(-1, self.embedding_dim)
        encoding_indices = self.get_code_indices(flat_x)
        quantized = self.quantize(encoding_indices).view_as(x_memout)
        quantized = quantized.permute(0, 3, 1, 2).contiguous()
        quantized = torch.unsqueeze(quantized, dim=0).expand(T, -1, -1, -1, -1)
        quantized = self.poisson(quantized)
        return quantized, encoding_indices

    def _forward_train(self, x: torch.Tensor):
        """Training branch (vae_model.py:42-85 with num_step := T): quantise the firing-rate / membrane feature with a
        straight-through estimator, regenerate spikes from the codes, and penalise both the feature distance (loss_1) and the
        distance between the post-synaptic potentials of the regenerated and the original spike trains (loss_2), each as
        codebook term + commitment_cost * commitment term."""
        T, beta = x.shape[0], self.commitment_cost
        feature = ((1 - self.alpha) * self.memout(x) + self.alpha * torch.sum(x, dim=0) / T).permute(0, 2, 3, 1).contiguous()
        idx = self.get_code_indices(feature.detach().reshape(-1, self.embedding_dim))          # argmin: no gradient
        codes = F.embedding(idx, self.embeddings.weight).view_as(feature)
        loss_1 = F.mse_loss(codes, feature.detach()) + beta * F.mse_loss(feature, codes.detach())
        straight_through = feature + (codes - feature).detach()
        regenerated = self.poisson(straight_through.permute(0, 3, 1, 2).contiguous().unsqueeze(0).repeat(T, 1, 1, 1, 1))
        psp = self.psp
        loss_2 = (torch.mean((psp(regenerated) - psp(x.detach())) ** 2)
                  + beta * torch.mean((psp(regenerated.detach()) - psp(x)) ** 2))
        return regenerated, loss_1 + loss_2

    @on_device_of
    def forward_with_loss(self, x: torch.Tensor):
        """(quantized, loss, indices) -- the 3-tuple named in BASELINE.json:north_star.
        train: the reference's training branch (vae_model.py:61-85), loss = loss_1 + loss_2 with gradients, indices
        from get_code_indices on the same feature; eval: the VQ objective value (no PSP term), for monitoring only."""
        if self.training:
            quantized, loss = self._forward_train(x)
            with torch.no_grad():
                T = x.shape[0]
                x_memout = (1 - self.alpha) * self.memout(x) + self.alpha * torch.sum(x, dim=0) / T
                flat = x_memout.permute(0, 2, 3, 1).reshape(-1, self.embedding_dim)
                idx = self.get_code_indices(flat)
            return quantized, loss, idx
        quantized, idx = self.forward(x)
        x_memout = self.feature(x)
        q = self.quantize(idx).view_as(x_memout)
        mse = torch.mean((q - x_memout) ** 2)
        return quantized, mse + self.commitment_cost * mse, idx

    @on_device_of
    def get_code_indices(self, flat_x: torch.Tensor) -> torch.Tensor:
        """argmin_k |z|^2 + |e_k|^2 - 2 z.e_k, first index on ties  (vae_model.py:87-95)."""
        if not flat_x.is_cuda:
            raise RuntimeError("get_code_indices needs CUDA tensors: there is no CPU path")
        z = flat_x.contiguous().float()
        idx = torch.empty(z.shape[0], dtype=torch.int64, device=z.device)
        check(lib().sd_vq_lookup(ptr(z), ptr(self.embeddings.weight.detach().contiguous()), ptr(idx), None, z.shape[0],
                                 self.embedding_dim, self.num_embeddings, stream_ptr()))
        return idx

    @on_device_of
    def quantize(self, encoding_indices: torch.Tensor) -> torch.Tensor:
        """Embedding rows for a tensor of indices: [...] -> [..., D]  (vae_model.py:97-99; called directly by
        R/main.py:264,389,424 with indices of shape [b, 7, 7])."""
        if not encoding_indices.is_cuda:
            raise RuntimeError("quantize needs CUDA tensors: there is no CPU path")
        if encoding_indices.numel() and (int(encoding_indices.min()) < 0 or int(encoding_indices.max()) >= self.num_embeddings):
            raise IndexError("index out of range in self")  # nn.Embedding's error
        shape = tuple(encoding_indices.shape)
        flat = encoding_indices.reshape(-1).contiguous().long()
        n = flat.numel()
        out = torch.empty((n, self.embedding_dim), dtype=torch.float32, device=flat.device)
        # gather with H = W = 1 writes [n, D, 1, 1] == [n, D]
        check(lib().sd_vq_gather(ptr(flat), ptr(self.embeddings.weight.detach().contiguous()), ptr(out), n,
                                 self.embedding_dim, 1, 1, self.num_embeddings, stream_ptr()))
        return out.view(*shape, self.embedding_dim)


class Encoder(nn.Module):
    """Encoder of VQ-VAE (vae_model.py:101-129)."""

    def __init__(self, in_dim=1, latent_dim=16):
        super().__init__()
        self.in_dim = in_dim
        self.latent_dim = latent_dim
        self.snn_convs = layer.SpikingSequential(
            layer.Conv2d(in_channels=in_dim, out_channels=32, kernel_size=3, stride=2, padding=1),
            layer.BatchNorm2d(32),
            neuron.LIFNode(surrogate_function=surrogate.ATan()),
            layer.Conv2d(in_channels=32, out_channels=64, kernel_size=3, stride=2, padding=1),
            layer.BatchNorm2d(64),
            neuron.LIFNode(surrogate_function=surrogate.ATan()),
            layer.Conv2d(in_channels=64, out_channels=latent_dim, kernel_size=1, stride=1, padding=0),
            layer.BatchNorm2d(latent_dim),
            neuron.LIFNode(surrogate_function=surrogate.ATan()),
        )

    def forward(self, x):
        return self.snn_convs(x)  # [t, b, c, h, w]


class Decoder(nn.Module):
    """Decoder of VQ-VAE (vae_model.py:131-159)."""

    def __init__(self, out_dim=1, latent_dim=16):
        super().__init__()
        self.out_dim = out_dim
        self.latent_dim = latent_dim
        self.snn_convs = layer.SpikingSequential(
            layer.ConvTranspose2d(in_channels=latent_dim, out_channels=64, kernel_size=3, stride=2, padding=1,
                                  output_padding=1),
            layer.BatchNorm2d(64),
            neuron.LIFNode(surrogate_function=surrogate.ATan()),
            layer.ConvTranspose2d(in_channels=64, out_channels=32, kernel_size=3, stride=2, padding=1,
                                  output_padding=1),
            layer.BatchNorm2d(32),
            neuron.LIFNode(surrogate_function=surrogate.ATan()),
            layer.ConvTranspose2d(in_channels=32, out_channels=out_dim, kernel_size=3, stride=1, padding=1,
                                  output_padding=0),
        )

    def forward(self, x):
        return self.snn_convs(x)  # [t, b, c, h, w]


class SNN_VQVAE(engine.PlanCacheMixin, nn.Module):
    """VQ-SVAE (vae_model.py:161-196)."""

    def __init__(self, in_dim, embedding_dim, num_embeddings, data_variance, commitment_cost=0.25, T: int = 16):
        super().__init__()
        self.in_dim = in_dim
        self.embedding_dim = embedding_dim
        self.num_embeddings = num_embeddings
        self.data_variance = data_variance
        self.T = T
        self.encoder = Encoder(in_dim, embedding_dim)
        self.vq_layer = VectorQuantizer(embedding_dim, num_embeddings, commitment_cost, T=T)
        self.decoder = Decoder(in_dim, embedding_dim)
        self.memout = MembraneOutputLayer(T)
        self._plans = {}

    def plan(self, T: int, B: int, H: int, W: int) -> "engine.VQVAEPlan":
        """Fully fused plan for a fixed shape (no fp32 round trips between stages; used by sampling/decoding)."""
        key = (T, B, H, W) + engine.module_cache_key(self)
        if self._plans.get("key") != key:
            self._plans = {"key": key, "plan": engine.VQVAEPlan(self, T, B, H, W)}
        return self._plans["plan"]

    @on_device_of
    def forward(self, x, image):
        """x: [T, B, C, H, W].  eval -> (e, x_recon, encoding_indices)   (vae_model.py:181-187);
        train -> (e_q_loss, recon_loss, real_recon_loss)                   (vae_model.py:189-196)."""
        if not self.training and isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 5 and self._all_lif_reset():
            return self._forward_fused(x)
        z = self.encoder(x)
        if not self.training:
            e, enco = self.vq_layer(z)
            x_recon = self.decoder(e)
            x_recon = self.memout(x_recon, apply_tanh=True)
            return e, x_recon, enco
        e, e_q_loss = self.vq_layer(z)
        x_recon = self.decoder(e)
        x_recon = torch.tanh(self.memout(x_recon))
        real_recon_loss = F.mse_loss(x_recon, image)
        recon_loss = real_recon_loss / self.data_variance
        return e_q_loss, recon_loss, real_recon_loss

    def _lif_nodes(self):
        """(plan layer name, LIF module) of every spiking layer of the eval forward, in execution order."""
        enc, gen, dec = self.encoder.snn_convs, self.vq_layer.poisson, self.decoder.snn_convs
        return (("e1", enc[2]), ("e2", enc[5]), ("e3", enc[8]), ("gen", gen[2]), ("d1", dec[2]), ("d2", dec[5]))

    def _all_lif_reset(self) -> bool:
        return all(isinstance(n, neuron.LIFNode) and n.memory_is_reset("v") for _, n in self._lif_nodes())

    @torch.no_grad()
    def _forward_fused(self, x):
        """Eval forward from freshly reset states as ONE fused chain (engine.VQVAEPlan) instead of module by module.
        The reference's state protocol is kept: every LIFNode.v holds the final membrane potential afterwards (built
        from the kernels' planar layout only if it is read before the next reset)."""
        T, B, _, H, W = x.shape
        plan = self.plan(T, B, H, W)
        e_stf, rec, idx = plan.forward(x.float(), capture_states=True)
        states = plan.states()
        for name, node in self._lif_nodes():
            node.v = base.LazyState(states[name])
        e = engine.stf_to_nchw(e_stf, T, B, self.embedding_dim, plan.h, plan.w)
        return e, rec.clone(), idx.clone()

    @on_device_of
    @torch.no_grad()
    def decode_indices(self, sample: torch.Tensor, T: int = None) -> torch.Tensor:
        """The caller-side decode of R/main.py:388-399 as one fused chain:
        quantize(sample) -> permute -> repeat(T) -> poisson -> decoder -> tanh(memout(.)).
        sample: int64 [b, h, w] (or [b,1,h,w]) -> fp32 [b, C, 4h, 4w] in (-1, 1)."""
        _no_training(self)
        T = self.T if T is None else T
        s = sample.reshape(sample.shape[0], *sample.shape[-2:])
        b, h, w = s.shape
        p = self.plan(T, b, 4 * h, 4 * w)
        return p.decode_indices(s.contiguous().long()).clone()
