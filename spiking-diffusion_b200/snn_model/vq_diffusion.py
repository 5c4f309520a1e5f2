"""Absorbing-state discrete diffusion and its spiking conv denoiser (mirrors R/snn_model/vq_diffusion.py:43-208).

Same class names, constructor arguments, attributes (``n_samples``, ``num_timesteps``, ``shape``, ``mask_id``,
``num_embeddings``) and ``state_dict`` keys (``conv{1..5}.{0,1}.*``, ``conv6.0.*``).  Additive differences:
* ``DummyModel(..., T=16)``: the reference hard-codes 16 timesteps (vq_diffusion.py:198,206);
* ``AbsorbingDiffusion(..., shape=(7,7), n_samples=16)``: the reference hard-codes the 7x7 grid, 49 steps and 16
  samples (vq_diffusion.py:47-51); they remain plain attributes that callers may overwrite, as in the reference;
* ``sample(..., seed=None)``: with ``seed=None`` the sampler consumes torch's CUDA generator exactly as the
  reference's two draws per step would (seed and offset are read from, and advanced on, the default generator),
  so ``torch.manual_seed(s); sample()`` reproduces the reference's stream.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib, engine
from .._lib import check, lib, on_device_of, ptr, stream_ptr
from ..activation_based import base
from ..activation_based import functional, layer, neuron, surrogate


def get_data_for_diff(train_loader, model, T: int = None):
    """Encode a dataset into code-index grids for diffusion training (mirrors vq_diffusion.py:23-36).

    Like the reference, the model is put in eval mode and is NOT reset between batches, so LIF membrane state leaks
    from one batch into the next; call ``functional.reset_net(model)`` per batch
    yourself if that is not wanted.  ``T`` defaults to the model's timestep count (the reference hard-codes 16)."""
    print('prepare data for train diffusion...')          # the reference's progress line (vq_diffusion.py:24)
    model.eval()
    steps = getattr(model, "T", 16) if T is None else T
    grids = []
    with torch.inference_mode():
        for batch, _labels in train_loader:
            frame = (batch - 0.5).cuda()                                    # ToTensor() range -> [-0.5, 0.5]
            codes = model(frame.unsqueeze(0).repeat(steps, 1, 1, 1, 1), frame)[2]   # direct encoding: the frame at every timestep
            grids.append(codes.reshape(frame.shape[0], frame.shape[-2] // 4, frame.shape[-1] // 4).cpu())
    return grids


def load_reference_state_dict(module, state_dict, strict: bool = True):
    """Load a reference checkpoint (R/main.py:199,286 ``torch.save(model.state_dict())``) into a module built with a
    different ``T``: the ``*.memout.coef`` buffers are T-bound ((16,1,1,1,1) in reference checkpoints, SURVEY.md
    finding 1) and are recomputed for this module's T instead of being copied."""
    own = module.state_dict()
    fixed = {}
    for k, v in state_dict.items():
        if k.endswith("memout.coef") and k in own and own[k].shape != v.shape:
            fixed[k] = own[k]
        else:
            fixed[k] = v
    return module.load_state_dict(fixed, strict=strict)


class DummyModel(engine.PlanCacheMixin, nn.Module):
    """6-layer spiking conv denoiser with one skip connection (vq_diffusion.py:150-208)."""

    def __init__(self, n_channel: int, num_embeddings, T: int = 16) -> None:
        super().__init__()
        self.num_embeddings = num_embeddings
        self.T = T
        def block(cin, cout):
            return layer.SpikingSequential(
                layer.Conv2d(in_channels=cin, out_channels=cout, kernel_size=3, stride=1, padding=1),
                layer.BatchNorm2d(cout),
                neuron.LIFNode(surrogate_function=surrogate.ATan()))
        self.conv1 = block(n_channel * 2, 64)
        self.conv2 = block(64, 128)
        self.conv3 = block(128, 256)
        self.conv4 = block(256, 512)
        self.conv5 = block(512, 256)
        self.conv6 = layer.SpikingSequential(layer.Conv2d(256 + 64, num_embeddings, 3, 1, 1))
        self._plans = {}
        # weight representation of the tcgen05 layers conv2..conv5: 3 = three int8 digits of a 22-bit fixed-point weight
        # (kind::i8, exact int32 accumulation; default, needs an even T, else falls back to 2), 2 = two fp16 terms
        # (kind::f16, 22 significant bits), 1 = one fp16 term (11 bits: NOT a parity configuration, experiments only)
        self.nsplit = 3

    def plan(self, b: int, h: int, w: int) -> "engine.DenoiserPlan":
        key = (self.T, b, h, w, self.nsplit) + engine.module_cache_key(self)
        if self._plans.get("key") != key:
            self._plans = {"key": key, "plan": engine.DenoiserPlan(self, self.T, b, h, w, nsplit=self.nsplit)}
        return self._plans["plan"]

    @on_device_of
    def forward(self, x, t) -> torch.Tensor:
        """x: [b, 1, h, w] float token ids, t: [b] long -> logits [b, K, h, w]  (vq_diffusion.py:189-208).

        The whole network runs as one fused chain, so the per-layer LIF states are consumed inside the kernels:
        this equals the reference whenever the net is reset between calls, which every reference call site does
        (vq_diffusion.py:129, R/main.py:249,391).  A second forward WITHOUT a reset would continue from the carried
        state in the reference; here it raises (the nodes are marked ``base.ConsumedState`` until reset)."""
        if not x.is_cuda:
            raise RuntimeError("DummyModel.forward needs CUDA tensors: there is no CPU path")
        if self.training:
            # training: the reference's layer-by-layer graph (vq_diffusion.py:195-206) on the autograd-capable kernels
            tt = torch.ones_like(x) * (t.unsqueeze(1).unsqueeze(2).unsqueeze(3))
            xin = torch.cat((x, tt), dim=1).unsqueeze(dim=0).repeat(self.T, 1, 1, 1, 1)
            x1 = self.conv1(xin)
            x5 = self.conv5(self.conv4(self.conv3(self.conv2(x1))))
            x6 = self.conv6(torch.cat((x5, x1), dim=2))
            return torch.sum(x6, dim=0) / self.T
        nodes = [(n, m) for n, m in self.named_modules() if isinstance(m, neuron.LIFNode)]
        for _, m in nodes:
            if not m.memory_is_reset("v"):
                raise RuntimeError("DummyModel.forward starts from reset LIF state; call functional.reset_net(model) first")
        b, _, h, w = x.shape
        logits = self.plan(b, h, w).run(x.float(), t)
        for n, m in nodes:
            m.v = base.ConsumedState(f"DummyModel.{n}")
        return logits.permute(0, 3, 1, 2).contiguous()


class Sampler(nn.Module):
    def __init__(self):
        super().__init__()


class AbsorbingDiffusion(engine.PlanCacheMixin, Sampler):
    def __init__(self, denoise_fn, mask_id, shape=(7, 7), n_samples: int = 16):
        super().__init__()
        self.num_classes = denoise_fn.num_embeddings
        self.shape = list(shape)
        self.num_timesteps = int(shape[0] * shape[1])
        self.mask_id = mask_id
        self._denoise_fn = denoise_fn
        self.n_samples = n_samples
        self.mask_schedule = "random"
        self.loss_type = "reweighted_elbo"
        self._plans = {}
        # larger n_samples are generated as consecutive chunks of this many images by one plan (activations of a chunk
        # stay close to the L2 size; the chunks are slices of ONE global Philox stream, so the result does not depend
        # on the chunk size)
        self.max_plan_batch = 4096

    def sample_time(self, b, device):
        t = torch.randint(1, self.num_timesteps + 1, (b,), device=device).long()
        pt = torch.ones_like(t).float() / self.num_timesteps
        return t, pt

    def q_sample(self, x_0, t):
        """Forward (masking) process of the absorbing diffusion (vq_diffusion.py:61-72): token (b, i) is replaced by the mask
        id with probability t_b / num_timesteps.  Returns (x_t, targets with -1 where nothing was masked, mask).  One uniform
        per token from torch's generator, drawn like the reference's ``rand_like`` (same shape, dtype, device)."""
        p_mask = t.to(torch.float32).view(-1, 1, 1, 1) / self.num_timesteps
        mask = torch.rand_like(x_0, dtype=torch.float32) < p_mask
        return x_0.masked_fill(mask, self.mask_id), x_0.masked_fill(~mask, -1), mask

    def _train_loss(self, x_0):
        """ELBO / re-weighted ELBO of the absorbing diffusion (vq_diffusion.py:75-101); x_0: [b, 1, h, w] token ids.  The
        cross entropy over the masked tokens (the others carry the ignore index) is summed per sample; 'elbo' divides it by
        t and by the probability 1/num_timesteps of having drawn that t, 'reweighted_elbo' weights it by 1 - t/num_timesteps;
        both are expressed in bits per token."""
        b = x_0.size(0)
        n_tok = self.shape[0] * self.shape[1]
        t, pt = self.sample_time(b, x_0.device)
        x_t, target, _ = self.q_sample(x_0=x_0, t=t)
        logits = self._denoise_fn(x_t, t=t)                                              # [b, K, h, w]
        nll = F.cross_entropy(logits.reshape(b, self.num_classes, n_tok), target.reshape(b, n_tok).long(),
                              ignore_index=-1, reduction='none').sum(1)
        bits = math.log(2) * x_0.shape[1:].numel()
        if self.loss_type == 'elbo':
            per_sample = nll / t / pt / bits
        elif self.loss_type == 'reweighted_elbo':
            per_sample = (1 - (t / self.num_timesteps)) * nll / bits
        else:
            raise ValueError
        return per_sample.mean()

    def train_iter(self, x):
        return {"loss": self._train_loss(x)}

    def plan(self, b: int, n_global=None, shard_base: int = 0) -> "engine.SamplerPlan":
        h, w = self.shape
        m = self._denoise_fn
        key = (b, h, w, m.T, m.nsplit, int(self.mask_id), n_global, shard_base) + engine.module_cache_key(m)
        if self._plans.get("key") != key:
            self._plans = {"key": key, "plan": engine.SamplerPlan(m, m.T, b, h, w, int(self.mask_id), n_global,
                                                                  shard_base, nsplit=m.nsplit)}
        return self._plans["plan"]

    def _tail_plan(self, n, n_global, shard_base):
        """Plan for the ragged last chunk of a chunked batch (kept beside the main plan, same invalidation key)."""
        key = ("tail", n, n_global, shard_base, self._plans.get("key"))
        if self._plans.get("tail_key") != key:
            m, (h, w) = self._denoise_fn, self.shape
            self._plans["tail_key"] = key
            self._plans["tail"] = engine.SamplerPlan(m, m.T, n, h, w, int(self.mask_id), n_global, shard_base,
                                                     nsplit=m.nsplit)
        return self._plans["tail"]

    @on_device_of
    @torch.no_grad()
    def sample(self, temp=1.0, sample_steps=None, seed=None, offset=0, n_global=None, shard_base: int = 0,
               x_init=None, unmasked_init=None, history=None):
        """Reverse process (vq_diffusion.py:103-142) -> x_t int64 [b, 1, h, w] with no mask tokens left.

        ``n_global`` / ``shard_base``: this call generates images [shard_base, shard_base + n_samples) of a global
        batch of ``n_global``; each shard evaluates the Philox values of its global element indices, so sharding the
        batch over GPUs reproduces the single-GPU stream (no collective on this path).
        ``history``: optional CUDA int64 [sample_steps, b*h*w] buffer that receives the token grid after every step
        (parity diagnostics; the loop then runs eagerly instead of as a CUDA graph)."""
        if self._denoise_fn.training:
            raise NotImplementedError("sample() runs the denoiser in eval mode; call denoise_fn.eval()")
        b = int(self.n_samples)
        if sample_steps is None:
            raise TypeError("unsupported operand type(s) for +: 'NoneType' and 'int'")  # the reference's failure mode
        n_global = b if n_global is None else int(n_global)
        chunk = int(self.max_plan_batch)
        if b > chunk and (x_init is not None or unmasked_init is not None or history is not None):
            raise ValueError(f"x_init / unmasked_init / history are per-plan buffers: use n_samples <= {chunk}")
        plan = self.plan(min(b, chunk), n_global, shard_base)
        if seed is None:
            gen = torch.cuda.default_generators[torch.cuda.current_device()]
            seed, offset = gen.initial_seed(), gen.get_offset()
            gen.set_offset(offset + plan.offset_advance(sample_steps))
        if b <= chunk:
            x_t = plan.sample(float(temp), int(sample_steps), int(seed), int(offset), x_init, unmasked_init,
                              history=history).clone()
        else:
            parts, lo = [], 0
            while lo < b:
                n = min(chunk, b - lo)
                p = plan if n == chunk else self._tail_plan(n, n_global, shard_base + lo)
                parts.append(p.sample(float(temp), int(sample_steps), int(seed), int(offset),
                                      extra_base=lo if n == chunk else 0).clone())
                lo += n
            x_t = torch.cat(parts)
        functional.reset_net(self._denoise_fn)
        return x_t
