"""MembraneOutputLayer and PSP (mirrors R/snn_model/snn_layers.py:6-41).

The reference hard-codes ``n_steps = 16`` (snn_layers.py:31); here the number of timesteps is a constructor
argument with the same default, so reference checkpoints (whose ``coef`` buffer has shape (16,1,1,1,1)) load
unchanged and T = 4 / 8 configurations are expressible.
"""
import ctypes

import torch
import torch.nn as nn

from .._lib import check, lib, on_device_of, ptr, stream_ptr


class MembraneOutputLayer(nn.Module):
    def __init__(self, n_steps: int = 16) -> None:
        super().__init__()
        arr = torch.arange(n_steps - 1, -1, -1)
        self.register_buffer("coef", torch.pow(0.8, arr)[:, None, None, None, None])  # (T,1,1,1,1)

    def coef_host(self, T: int):
        c = self.coef.detach().reshape(-1).cpu().tolist()
        if len(c) != T:
            # the reference raises a broadcasting RuntimeError here (SURVEY.md finding 1)
            raise RuntimeError(f"The size of tensor a ({T}) must match the size of tensor b ({len(c)}) at "
                               "non-singleton dimension 0")
        return (ctypes.c_float * T)(*c)

    @on_device_of
    def forward(self, x: torch.Tensor, apply_tanh: bool = False) -> torch.Tensor:
        """x: (T, N, C, H, W) -> sum_t coef[t] * x[t]."""
        if not x.is_cuda:
            raise RuntimeError("MembraneOutputLayer.forward needs CUDA tensors: there is no CPU path")
        T = x.shape[0]
        if torch.is_grad_enabled() and x.requires_grad:
            # training: a T-term weighted sum; kept as tensor algebra so that autograd provides the backward
            out = torch.sum(x * self.coef, dim=0)
            return torch.tanh(out) if apply_tanh else out
        coef = self.coef_host(T)
        xc = x.contiguous().float()
        out = torch.empty(xc.shape[1:], dtype=torch.float32, device=x.device)
        check(lib().sd_memout(ptr(xc), ptr(out), ctypes.cast(coef, ctypes.c_void_p), T, out.numel(), int(apply_tanh),
                              stream_ptr()))
        return out


class PSP(nn.Module):
    """Post-synaptic potential low-pass, tau_s = 2 (snn_layers.py:6-26).  Only used by the training loss
    (R/snn_model/vae_model.py:81-82), which is outside this round's scope; kept as plain tensor algebra."""

    def __init__(self):
        super().__init__()
        self.tau_s = 2

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        syn = torch.zeros_like(inputs[0])
        out = []
        for t in range(inputs.shape[0]):
            syn = syn + (inputs[t] - syn) / self.tau_s
            out.append(syn)
        return torch.stack(out)
