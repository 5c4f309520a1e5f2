"""LIF neuron backed by the sm_100a kernel ``sd_lif_forward`` (mirrors SJ/activation_based/neuron.py:23-263,603-1010).

``LIFNode(step_mode='m').forward(x_seq[T, N, ...])`` returns a spike tensor of the same shape and dtype and keeps
the membrane potential ``v`` between calls until ``reset()`` -- the contract of the reference's eval path
(neuron.py:971-1010 -> :799-809).  ``MultiStepLIFNode`` is the older SpikingJelly name for the same thing with
``step_mode='m'`` preset.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .. import _lib
from .._lib import check, lib, ptr, stream_ptr
from . import base, surrogate


class _LIFFunction(torch.autograd.Function):
    """Multi-step LIF with surrogate-gradient BPTT: forward = sd_lif_forward (saving the pre-fire potential),
    backward = sd_lif_backward.  Inputs: x_seq [T, ...], v_init [...]; outputs: spike_seq, v_last."""

    @staticmethod
    def forward(ctx, x_seq, v_init, tau, v_th, v_reset, decay_input, detach_reset, alpha):
        x_seq = x_seq.contiguous()
        v = v_init.contiguous().clone()
        spikes = torch.empty_like(x_seq)
        h_seq = torch.empty_like(x_seq)
        T, N = x_seq.shape[0], x_seq[0].numel()
        check(lib().sd_lif_forward(ptr(x_seq), ptr(v), ptr(spikes), ptr(h_seq), T, N, float(tau), float(v_th),
                                   0.0 if v_reset is None else float(v_reset), int(v_reset is not None), int(decay_input),
                                   stream_ptr()))
        ctx.save_for_backward(h_seq)
        ctx.cfg = (float(tau), float(v_th), v_reset, bool(decay_input), bool(detach_reset), float(alpha))
        return spikes, v

    @staticmethod
    def backward(ctx, grad_spikes, grad_v_last):
        h_seq, = ctx.saved_tensors
        tau, v_th, v_reset, decay_input, detach_reset, alpha = ctx.cfg
        T, N = h_seq.shape[0], h_seq[0].numel()
        gs = grad_spikes.contiguous().float() if grad_spikes is not None else torch.zeros_like(h_seq)
        gv = grad_v_last.contiguous().float() if grad_v_last is not None else None
        grad_x = torch.empty_like(h_seq)
        grad_v0 = torch.empty_like(h_seq[0])
        check(lib().sd_lif_backward(ptr(gs), ptr(gv), ptr(h_seq), ptr(grad_x), ptr(grad_v0), T, N, tau, v_th,
                                    0.0 if v_reset is None else float(v_reset), int(v_reset is not None), int(decay_input),
                                    int(detach_reset), alpha, stream_ptr()))
        return grad_x, grad_v0, None, None, None, None, None, None


class BaseNode(base.MemoryModule):
    def __init__(self, v_threshold: float = 1.0, v_reset: Optional[float] = 0.0,
                 surrogate_function: Callable = surrogate.Sigmoid(), detach_reset: bool = False, step_mode="s",
                 backend="torch", store_v_seq: bool = False):
        # same argument checks as SJ/activation_based/neuron.py:89-94
        assert isinstance(v_reset, float) or v_reset is None
        assert isinstance(v_threshold, float)
        assert isinstance(detach_reset, bool)
        super().__init__()
        self.register_memory("v", 0.0 if v_reset is None else v_reset)
        self.v_threshold = v_threshold
        self.v_reset = v_reset
        self.detach_reset = detach_reset
        self.surrogate_function = surrogate_function
        self.step_mode = step_mode
        self.backend = backend
        self.store_v_seq = store_v_seq

    @property
    def store_v_seq(self):
        return self._store_v_seq

    @store_v_seq.setter
    def store_v_seq(self, value: bool):
        self._store_v_seq = value
        if value and not hasattr(self, "v_seq"):
            self.register_memory("v_seq", None)

    def v_float_to_tensor(self, x: torch.Tensor):
        # neuron.py:260-263: the first call materialises v = full_like(x, v_reset)
        if isinstance(self.v, float):
            self.v = torch.full_like(x.data, self.v)

    def extra_repr(self):
        return (f"v_threshold={self.v_threshold}, v_reset={self.v_reset}, detach_reset={self.detach_reset}, "
                f"step_mode={self.step_mode}, backend={self.backend}")


class LIFNode(BaseNode):
    def __init__(self, tau: float = 2.0, decay_input: bool = True, v_threshold: float = 1.0,
                 v_reset: Optional[float] = 0.0, surrogate_function: Callable = surrogate.Sigmoid(),
                 detach_reset: bool = False, step_mode="s", backend="torch", store_v_seq: bool = False):
        assert isinstance(tau, float) and tau > 1.0  # neuron.py:707
        super().__init__(v_threshold, v_reset, surrogate_function, detach_reset, step_mode, backend, store_v_seq)
        self.tau = tau
        self.decay_input = decay_input

    @property
    def supported_backends(self):
        # The reference lists ('torch',) / ('torch', 'cupy'); both names are accepted so that code calling
        # functional.set_backend keeps working.  Every name runs the same sm_100a kernel: there is no dispatch.
        return ("torch", "cupy", "b200")

    def extra_repr(self):
        return super().extra_repr() + f", tau={self.tau}"

    def _run(self, x_seq: torch.Tensor) -> torch.Tensor:
        if not x_seq.is_cuda:
            raise RuntimeError("LIFNode.forward needs a CUDA tensor: spiking_diffusion_b200 has no CPU path")
        if x_seq.dtype != torch.float32:
            raise NotImplementedError(f"LIFNode kernel is fp32 only, got {x_seq.dtype}")
        if x_seq.shape[0] == 0:
            return torch.empty_like(x_seq)
        x_seq = x_seq.contiguous()
        self.v_float_to_tensor(x_seq[0])
        if self.v.shape != x_seq.shape[1:]:
            raise ValueError(f"membrane state shape {tuple(self.v.shape)} does not match input {tuple(x_seq.shape[1:])}; "
                             "call reset() between inputs of different shape")
        if self.training and torch.is_grad_enabled() and (x_seq.requires_grad or self.v.requires_grad):
            # training branch (neuron.py:244-258): same spikes as the eval kernel, gradients by surrogate BPTT
            sf = self.surrogate_function
            if not isinstance(sf, surrogate.ATan) or not getattr(sf, "spiking", True):
                raise NotImplementedError("surrogate-gradient BPTT is implemented for the spiking ATan surrogate only")
            spikes, v = _LIFFunction.apply(x_seq, self.v, self.tau, self.v_threshold, self.v_reset, self.decay_input,
                                           self.detach_reset, sf.alpha)
            self.v = v
            if self.store_v_seq:
                raise NotImplementedError("store_v_seq is not available on the autograd path")
            spikes._sd_is_spikes = True      # lets the next layer.Conv2d take the tensor-core forward (layer._ConvFn)
            return spikes
        v = self.v.contiguous()
        spikes = torch.empty_like(x_seq)
        h_seq = torch.empty_like(x_seq) if self.store_v_seq else None
        T, N = x_seq.shape[0], x_seq[0].numel()
        check(lib().sd_lif_forward(ptr(x_seq), ptr(v), ptr(spikes), ptr(h_seq), T, N, float(self.tau),
                                   float(self.v_threshold), 0.0 if self.v_reset is None else float(self.v_reset),
                                   int(self.v_reset is not None), int(self.decay_input), stream_ptr()))
        self.v = v
        if self.store_v_seq:
            # v_seq[t] is the potential AFTER reset at step t (neuron.py:811-824)
            if self.v_reset is None:
                self.v_seq = h_seq - spikes * self.v_threshold
            else:
                self.v_seq = self.v_reset * spikes + (1.0 - spikes) * h_seq
        return spikes

    def multi_step_forward(self, x_seq: torch.Tensor):
        return self._run(x_seq)

    def single_step_forward(self, x: torch.Tensor):
        return self._run(x.unsqueeze(0))[0]


class MultiStepLIFNode(LIFNode):
    """``clock_driven``-era name: a LIFNode whose forward takes ``x_seq[T, N, C, H, W]``."""

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("step_mode", "m")
        super().__init__(*args, **kwargs)
