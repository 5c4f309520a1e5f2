"""Step-mode layer wrappers and the fused container.

``Conv2d`` / ``ConvTranspose2d`` / ``BatchNorm2d`` keep the constructor signatures, parameter names and the
``step_mode`` protocol of SJ/activation_based/layer.py:125-173, 276-325, 423-465, so reference ``state_dict``s load
unchanged.  Called on their own they run a single un-fused sm_100a kernel each.  ``SpikingSequential`` is the
``nn.Sequential`` the models are built from: it recognises conv -> BN -> LIF triplets (and a bare trailing conv)
and runs each as ONE fused kernel, with spike tensors staying in the packed STF layout between stages.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib, engine
from .._lib import check, lib, ptr, stream_ptr
from . import base, neuron


def _check_5d(x: torch.Tensor):
    # same condition, message shape and exception type as SJ/activation_based/layer.py:169-170
    if x.dim() != 5:
        raise ValueError(f"expected x with shape [T, N, C, H, W], but got x with shape {x.shape}!")


class _ConvSpec:
    """Minimal stand-in for an nn.Conv2d / nn.ConvTranspose2d (what engine.FusedLayer reads), used to run the adjoint
    convolution of a layer for its input gradient."""

    def __init__(self, weight, transposed, kernel_size, stride, padding, output_padding):
        self.weight, self.bias, self.transposed = weight, None, transposed
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.output_padding, self.dilation, self.groups = output_padding, (1, 1), 1


class _ConvFn(torch.autograd.Function):
    """Conv2d / ConvTranspose2d on [T, N, C, H, W] with gradients: forward and input gradient through the CUDA-core
    conv kernel (the input gradient of a convolution is the adjoint convolution with the same weights), weight /
    bias gradient through sd_conv_wgrad.  Mirrors torch autograd over F.conv2d / F.conv_transpose2d
    (SJ/activation_based/layer.py:164-173, 316-325)."""

    @staticmethod
    def _tc_forward(x5, mod):
        """Forward on the tcgen05 kind::i8 kernel for a spike input (tagged by LIFNode's training branch) when the layer is
        one the kernel takes (3x3, stride 1, pad 1, C_in % 32 == 0, C_out % 16 == 0, even T <= 16, grid width <= 62):
        u8 spikes x three exact int8 weight digits (22-bit weights: relative error <= 2^-22 per weight, far inside the 1e-5
        bar of tests/test_gpu_training.py), int32 accumulation.  Returns None when not applicable (SD_TRAIN_TC=0: never)."""
        if mod.transposed or os.environ.get("SD_TRAIN_TC", "1") == "0":
            return None
        T, B, _, H, W = x5.shape
        try:
            plan = engine.FusedLayer(mod, None, None, T=T, B=B, H_in=H, W_in=W, in_kind=_lib.IN_STF,
                                     out_kind=_lib.OUT_CURRENT_SEQ, impl="tc", nsplit=3)
        except ValueError:
            return None
        cur = plan.run(engine.stf8_from_nchw(x5), plan.alloc_out())
        return engine.currents_to_nchw(cur, T, B, plan.C_out, plan.H_out, plan.W_out)

    @staticmethod
    def forward(ctx, x5, weight, bias, mod):
        T, B, _, H, W = x5.shape
        spikes_in = bool(getattr(x5, "_sd_is_spikes", False))
        x5 = x5.contiguous().float()
        # the CUDA-core plan also provides the descriptor the backward pass works from
        plan = engine.FusedLayer(mod, None, None, T=T, B=B, H_in=H, W_in=W, in_kind=_lib.IN_REAL_SEQ,
                                 out_kind=_lib.OUT_REAL_SEQ, impl="simt")
        y = _ConvFn._tc_forward(x5, mod) if spikes_in else None
        if y is None:
            y = plan.run(x5, plan.alloc_out())
        ctx.save_for_backward(x5, weight)
        ctx.meta = (plan.desc, bool(mod.transposed), tuple(mod.kernel_size), tuple(mod.stride), tuple(mod.padding),
                    bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x5, weight = ctx.saved_tensors
        d, transposed, ks, stride, padding, has_bias = ctx.meta
        gy = gy.contiguous().float()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if transposed:   # adjoint of conv_transpose2d(x, w) is conv2d(gy, w): w [C_in, C_out, k, k] read as [out, in, k, k]
                spec = _ConvSpec(weight.detach(), False, ks, stride, padding, (0, 0))
            else:            # adjoint of conv2d(x, w) is conv_transpose2d(gy, w) sized back to the input
                op_h = d.H_in - ((d.H_out - 1) * stride[0] - 2 * padding[0] + ks[0])
                op_w = d.W_in - ((d.W_out - 1) * stride[1] - 2 * padding[1] + ks[1])
                spec = _ConvSpec(weight.detach(), True, ks, stride, padding, (op_h, op_w))
            adj = engine.FusedLayer(spec, None, None, T=d.T, B=d.B, H_in=d.H_out, W_in=d.W_out, in_kind=_lib.IN_REAL_SEQ,
                                    out_kind=_lib.OUT_REAL_SEQ, impl="simt")
            gx = adj.run(gy, adj.alloc_out())
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            gw = torch.empty_like(weight, dtype=torch.float32)
            gb = torch.empty(d.C_out, dtype=torch.float32, device=gy.device) if has_bias else None
            ws = torch.empty(lib().sd_conv_wgrad_workspace_bytes(ctypes.byref(d)), dtype=torch.uint8, device=gy.device)
            check(lib().sd_conv_wgrad(ctypes.byref(d), ptr(x5), ptr(gy), ptr(gw), ptr(gb), ptr(ws), stream_ptr()))
        return gx, gw, gb, None


class _BNFn(torch.autograd.Function):
    """Train-mode BatchNorm2d over [n_outer, C, H, W] (batch statistics), forward and backward on our kernels."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = x.contiguous().float()
        n_outer, C = x.shape[0], x.shape[1]
        hw = x.shape[2] * x.shape[3]
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        var = torch.empty(C, dtype=torch.float32, device=x.device)
        check(lib().sd_bn_train_forward(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(var), n_outer, C, hw,
                                        float(eps), stream_ptr()))
        ctx.save_for_backward(x, mean, var, gamma if gamma is not None else torch.empty(0, device=x.device))
        ctx.meta = (float(eps), gamma is not None, beta is not None)
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, gy, _gm, _gv):
        x, mean, var, gamma = ctx.saved_tensors
        eps, has_g, has_b = ctx.meta
        gy = gy.contiguous().float()
        n_outer, C = x.shape[0], x.shape[1]
        hw = x.shape[2] * x.shape[3]
        gx = torch.empty_like(x)
        gg = torch.empty(C, dtype=torch.float32, device=x.device)
        gb = torch.empty(C, dtype=torch.float32, device=x.device)
        check(lib().sd_bn_backward(ptr(x), ptr(gy), ptr(mean), ptr(var), ptr(gamma) if has_g else None, ptr(gx), ptr(gg),
                                   ptr(gb), n_outer, C, hw, eps, stream_ptr()))
        return gx, (gg if has_g else None), (gb if has_b else None), None


class _SyncBNFn(torch.autograd.Function):
    """Train-mode BatchNorm2d whose batch statistics span every rank of a process group (the reference would wrap its
    model in torch.nn.SyncBatchNorm; SURVEY.md 8(e): train-mode BN is the one batch-coupled op of the training path).
    Forward: local (count, mean, M2) per channel -> all_gather over NCCL -> parallel-variance combination -> affine.
    Backward: local sums of gy and gy * xhat -> all_reduce -> input gradient; the gamma / beta gradients stay local
    (DistributedDataParallel averages parameter gradients itself, as with torch's SyncBatchNorm)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, group):
        import torch.distributed as dist
        x = x.contiguous().float()
        n_outer, C = x.shape[0], x.shape[1]
        hw = x.shape[2] * x.shape[3]
        stats = torch.empty((3, C), dtype=torch.float32, device=x.device)
        check(lib().sd_bn_local_stats(ptr(x), ptr(stats[0]), ptr(stats[1]), n_outer, C, hw, stream_ptr()))
        stats[2].fill_(float(n_outer * hw))
        world = dist.get_world_size(group)
        gathered = torch.empty((world, 3, C), dtype=torch.float32, device=x.device)
        dist.all_gather_into_tensor(gathered, stats, group=group)
        cnt = gathered[:, 2]                                   # [world, C]
        total = cnt.sum(0)
        mean = (gathered[:, 0] * cnt).sum(0) / total
        m2 = gathered[:, 1].sum(0) + (cnt * (gathered[:, 0] - mean) ** 2).sum(0)
        var = m2 / total                                       # biased, like F.batch_norm uses for normalisation
        invstd = torch.rsqrt(var + eps)
        scale = (gamma if gamma is not None else torch.ones_like(mean)) * invstd
        shift = (beta if beta is not None else torch.zeros_like(mean)) - mean * scale
        y = torch.empty_like(x)
        check(lib().sd_channel_affine(ptr(x), ptr(scale.contiguous()), ptr(shift.contiguous()), ptr(y), n_outer, C, hw,
                                      stream_ptr()))
        mean, var = mean.contiguous(), var.contiguous()
        ctx.save_for_backward(x, mean, var, gamma if gamma is not None else torch.empty(0, device=x.device), total)
        ctx.meta = (float(eps), gamma is not None, beta is not None, group)
        ctx.mark_non_differentiable(mean, var, total)
        return y, mean, var, total

    @staticmethod
    def backward(ctx, gy, _gm, _gv, _gt):
        import torch.distributed as dist
        x, mean, var, gamma, total = ctx.saved_tensors
        eps, has_g, has_b, group = ctx.meta
        gy = gy.contiguous().float()
        n_outer, C = x.shape[0], x.shape[1]
        hw = x.shape[2] * x.shape[3]
        sums = torch.empty((2, C), dtype=torch.float32, device=x.device)
        check(lib().sd_bn_backward_reduce(ptr(x), ptr(gy), ptr(mean), ptr(var), ptr(sums[0]), ptr(sums[1]), n_outer, C, hw,
                                          eps, stream_ptr()))
        local = sums.clone()
        dist.all_reduce(sums, group=group)
        means = (sums / total).contiguous()
        gx = torch.empty_like(x)
        check(lib().sd_bn_backward_apply(ptr(x), ptr(gy), ptr(mean), ptr(var), ptr(gamma) if has_g else None, ptr(means[0]),
                                         ptr(means[1]), ptr(gx), n_outer, C, hw, eps, stream_ptr()))
        return gx, (local[1] if has_g else None), (local[0] if has_b else None), None, None


class _ConvMixin(base.StepModule):
    def _plan(self, T, B, H, W) -> engine.FusedLayer:
        key = (T, B, H, W, self.weight.data_ptr(), self.weight._version,
               None if self.bias is None else self.bias._version)
        if getattr(self, "_plan_key", None) != key:
            self._plan_obj = engine.FusedLayer(self, None, None, T=T, B=B, H_in=H, W_in=W,
                                               in_kind=_lib.IN_REAL_SEQ, out_kind=_lib.OUT_REAL_SEQ, impl="simt")
            self._plan_key = key
        return self._plan_obj

    def _forward_any(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("layer forward needs CUDA tensors: spiking_diffusion_b200 has no CPU path")
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad)
        if self.step_mode == "s":
            if x.dim() != 4:
                raise ValueError(f"expected x with shape [N, C, H, W], but got x with shape {x.shape}!")
            x5 = x.unsqueeze(0)
        else:
            _check_5d(x)
            x5 = x
        T, B, _, H, W = x5.shape
        if needs_grad:
            out = _ConvFn.apply(x5, self.weight, self.bias, self)
        else:
            plan = self._plan(T, B, H, W)
            out = plan.alloc_out()
            plan.run(x5.contiguous().float(), out)
        return out[0] if self.step_mode == "s" else out


class Conv2d(nn.Conv2d, _ConvMixin):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode="zeros", step_mode="s"):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode)
        self.step_mode = step_mode

    def extra_repr(self):
        return super().extra_repr() + f", step_mode={self.step_mode}"

    @_lib.on_device_of
    def forward(self, x: torch.Tensor):
        return self._forward_any(x)


class ConvTranspose2d(nn.ConvTranspose2d, _ConvMixin):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, groups=1,
                 bias=True, dilation=1, padding_mode="zeros", step_mode="s"):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, output_padding, groups, bias,
                         dilation, padding_mode)
        self.step_mode = step_mode

    def extra_repr(self):
        return super().extra_repr() + f", step_mode={self.step_mode}"

    @_lib.on_device_of
    def forward(self, x: torch.Tensor):
        return self._forward_any(x)


class BatchNorm2d(nn.BatchNorm2d, base.StepModule):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True, step_mode="s"):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)
        self.step_mode = step_mode

    def extra_repr(self):
        return super().extra_repr() + f", step_mode={self.step_mode}"

    @_lib.on_device_of
    def forward(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("layer forward needs CUDA tensors: spiking_diffusion_b200 has no CPU path")
        if self.step_mode == "m":
            _check_5d(x)
        elif x.dim() != 4:
            raise ValueError(f"expected x with shape [N, C, H, W], but got x with shape {x.shape}!")
        if self.training or not self.track_running_stats:
            # batch statistics over T*N*H*W (functional.seq_to_ann_forward flattens T into the batch, layer.py:458-465)
            x4 = x.flatten(0, 1) if x.dim() == 5 else x
            group = getattr(self, "sync_group", None)
            if group is not None and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
                y, mean, var, total = _SyncBNFn.apply(x4, self.weight, self.bias, self.eps, group)
                unbias = total / (total - 1).clamp(min=1)     # global element count per channel, kept on the device
            else:
                y, mean, var = _BNFn.apply(x4, self.weight, self.bias, self.eps)
                n = x4.numel() // x4.shape[1]
                unbias = n / max(n - 1, 1)
            if self.track_running_stats:
                with torch.no_grad():   # F.batch_norm's running update: momentum, unbiased variance
                    self.num_batches_tracked += 1
                    mom = self.momentum if self.momentum is not None else 1.0 / float(self.num_batches_tracked)
                    self.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
                    self.running_var.mul_(1 - mom).add_(var * unbias, alpha=mom)
            return y.view(x.shape)
        scale, shift = engine.fold_bn(None, self.num_features, self, x.device)
        xc = x.contiguous().float()
        out = torch.empty_like(xc)
        C = xc.shape[-3]
        hw = xc.shape[-1] * xc.shape[-2]
        check(lib().sd_channel_affine(ptr(xc), ptr(scale), ptr(shift), ptr(out), xc.numel() // (C * hw), C, hw,
                                      stream_ptr()))
        return out


# --------------------------------------------------------------------------------------------------
class SpikingSequential(nn.Sequential):
    """``nn.Sequential`` of (conv, BN, LIF) triplets executed as fused kernels.

    Same child names as a plain ``nn.Sequential`` -> same ``state_dict`` keys as the reference
    (e.g. ``encoder.snn_convs.{0,1,3,4,6,7}.*``).  Input and output use the reference's tensor format, fp32
    ``[T, N, C, H, W]``; in between, spikes stay in the STF layout.  LIF state follows the reference's
    protocol: each ``LIFNode.v`` persists between calls until ``reset()``.
    """

    def _stages(self) -> List[Tuple[nn.Module, Optional[nn.Module], Optional[nn.Module]]]:
        mods = list(self.children())
        out, i = [], 0
        while i < len(mods):
            conv = mods[i]
            if not isinstance(conv, (nn.Conv2d, nn.ConvTranspose2d)):
                raise TypeError(f"SpikingSequential expects conv[-BN][-LIF] groups, got {type(conv).__name__} at {i}")
            bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d) else None
            j = i + 1 + (bn is not None)
            lif = mods[j] if j < len(mods) and isinstance(mods[j], neuron.LIFNode) else None
            i = j + (lif is not None)
            out.append((conv, bn, lif))
        return out

    def _build(self, T, B, H, W, device):
        stages = self._stages()
        plans, h, w = [], H, W
        for k, (conv, bn, lif) in enumerate(stages):
            in_kind = _lib.IN_REAL_SEQ if k == 0 else _lib.IN_STF
            if lif is not None:
                out_kind = _lib.OUT_LIF
            else:
                if k != len(stages) - 1:
                    raise TypeError("a conv without LIF is only supported as the last stage")
                out_kind = _lib.OUT_REAL_SEQ
            if bn is not None and (bn.training or not bn.track_running_stats):
                raise NotImplementedError("train-mode BatchNorm (batch statistics) is not implemented in this round")
            fl = None
            if (k > 0 and out_kind == _lib.OUT_LIF and engine._UpsampledConvT.eligible(conv) and 2 * w + 2 <= 64
                    and os.environ.get("SD_DECODER_TC", "1") != "0"):
                # stride-2 transposed conv on spikes: the tcgen05 kernel on the zero-inserted upsampled input
                cand = engine.FusedLayer(engine._UpsampledConvT(conv), bn, lif, T=T, B=B, H_in=2 * h, W_in=2 * w,
                                         in_kind=in_kind, out_kind=out_kind, impl="auto")
                if cand.impl == "tc":
                    fl = cand
                    fl.upsample_src = (conv.in_channels, h, w)
                    fl.upsample_buf = engine.stf_empty(T, B, conv.in_channels, 2 * h, 2 * w, device)
            if fl is None:
                fl = engine.FusedLayer(conv, bn, lif, T=T, B=B, H_in=h, W_in=w, in_kind=in_kind, out_kind=out_kind,
                                       impl="simt" if k == 0 else "auto")
                fl.upsample_src = None
            plans.append(fl)
            h, w = fl.H_out, fl.W_out
        return stages, plans

    def invalidate_plans(self) -> None:
        """Drop the cached fused plan (packed weights, folded BN).  The cache key covers parameter / buffer versions and
        the LIF / BN hyper-parameters, but NOT in-place edits through ``.data`` (they do not bump ``_version``)."""
        self._key = None

    def __getstate__(self):
        # plans hold ctypes pointers / device buffers: never pickled or deep-copied, rebuilt on the next forward
        state = dict(self.__dict__)
        for k in ("_key", "_stage_list", "_plans", "_bufs"):
            state.pop(k, None)
        return state

    @_lib.on_device_of
    def forward(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("SpikingSequential.forward needs CUDA tensors: spiking_diffusion_b200 has no CPU path")
        _check_5d(x)
        if self.training:
            # training: layer by layer through the autograd-capable kernels (conv, train-mode BN, surrogate LIF), exactly
            # the reference's nn.Sequential; the fused inference kernels fold BN running statistics and have no backward
            for m in self:
                x = m(x)
            return x
        T, B, _, H, W = x.shape
        key = (T, B, H, W) + engine.module_cache_key(self)
        if getattr(self, "_key", None) != key:
            self._stage_list, self._plans = self._build(T, B, H, W, x.device)
            self._bufs = [p.alloc_out() for p in self._plans]
            self._key = key
        cur = x.contiguous().float()
        for (conv, bn, lif), plan, buf in zip(self._stage_list, self._plans, self._bufs):
            v_planar = None
            if lif is not None:
                d = plan.desc
                v_planar = plan.alloc_state()
                if not lif.memory_is_reset("v") and isinstance(lif.v, torch.Tensor):  # continue from the stored state (neuron.py:972-1010)
                    check(lib().sd_state_convert(ptr(lif.v.contiguous().float()), ptr(v_planar), d.B, d.C_out, d.H_out,
                                                 d.W_out, 1, stream_ptr()))
            out = buf if plan.desc.out_kind == _lib.OUT_LIF else plan.alloc_out()
            if plan.upsample_src is not None:
                c_in, h_in, w_in = plan.upsample_src
                check(lib().sd_stf_upsample2x(ptr(cur), ptr(plan.upsample_buf), plan.T, plan.B, c_in, h_in, w_in,
                                              stream_ptr()))
                cur = plan.upsample_buf
            plan.run(cur, out, v=v_planar)
            if lif is not None:
                # the final membrane potential stays in the kernels' planar layout; LIFNode.v is built from it only if
                # it is read before the next reset (functional.reset_net follows every forward in the reference)
                def to_reference_layout(buf=v_planar, d=plan.desc):
                    v_new = torch.empty((d.B, d.C_out, d.H_out, d.W_out), dtype=torch.float32, device=buf.device)
                    check(lib().sd_state_convert(ptr(buf), ptr(v_new), d.B, d.C_out, d.H_out, d.W_out, 0, stream_ptr()))
                    return v_new
                lif.v = base.LazyState(to_reference_layout)
            cur = out
        last = self._plans[-1].desc
        if last.out_kind == _lib.OUT_LIF:
            return engine.stf_to_nchw(cur, last.T, last.B, last.C_out, last.H_out, last.W_out)
        return cur
