"""Surrogate spike functions (mirrors SJ/activation_based/surrogate.py:12-50, 663-756).

The forward pass is the Heaviside step ``(x >= 0)``; on the fused CUDA path it is the ``h >= v_th`` comparison in
the kernel epilogue, and these modules only carry the hyper-parameters (``alpha``, ``spiking``).  The backward
formula of ATan is kept for the training row of SURVEY.md section 8(f).
"""
import math

import torch
import torch.nn as nn


def heaviside(x: torch.Tensor) -> torch.Tensor:
    return (x >= 0).to(x)


class _ATanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha):
        if x.requires_grad:
            ctx.save_for_backward(x)
            ctx.alpha = alpha
        return heaviside(x)

    @staticmethod
    def backward(ctx, grad_output):
        # alpha / 2 / (1 + (pi/2 * alpha * x)^2) * grad      (surrogate.py:663-665)
        x, = ctx.saved_tensors
        return ctx.alpha / 2 / (1 + (math.pi / 2 * ctx.alpha * x).pow_(2)) * grad_output, None


class SurrogateFunctionBase(nn.Module):
    def __init__(self, alpha, spiking=True):
        super().__init__()
        self.spiking = spiking
        self.alpha = alpha

    def extra_repr(self):
        return f"alpha={self.alpha}, spiking={self.spiking}"


class ATan(SurrogateFunctionBase):
    def __init__(self, alpha=2.0, spiking=True):
        super().__init__(alpha, spiking)

    def forward(self, x: torch.Tensor):
        if self.spiking:
            return _ATanFn.apply(x, self.alpha)
        return (math.pi / 2 * self.alpha * x).atan_() / math.pi + 0.5


class Sigmoid(SurrogateFunctionBase):
    def __init__(self, alpha=4.0, spiking=True):
        super().__init__(alpha, spiking)

    def forward(self, x: torch.Tensor):
        if self.spiking:
            return heaviside(x)
        return (x * self.alpha).sigmoid()
