"""Fused execution plans over the C-ABI (libsd_b200.so).

A *plan* owns, for one fixed shape, the pre-packed weights, folded BatchNorm affine and the STF activation
buffers of a chain of conv -> BN -> LIF layers, and enqueues one C-ABI call per layer on the current CUDA
stream.  PyTorch is used for device memory and streams only.

  FusedLayer     one conv[-BN][-LIF] stage (SIMT or tcgen05 implementation)
  DenoiserPlan   DummyModel.forward            (R/snn_model/vq_diffusion.py:189-208)
  SamplerPlan    AbsorbingDiffusion.sample     (R/snn_model/vq_diffusion.py:103-142)
  VQVAEPlan      encoder / quantiser / generator / decoder of SNN_VQVAE (R/snn_model/vae_model.py:101-196)
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import ConvArgs, ConvDesc, check, lib, ptr, stream_ptr

def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: spiking_diffusion_b200 has no CPU path "
                           "(the CPU oracle lives under oracle/ and is test infrastructure only)")


def module_cache_key(module: torch.nn.Module) -> tuple:
    """Everything a cached plan of ``module`` depends on besides the shape: parameter and buffer versions (bumped by
    optimizer steps, ``load_state_dict`` and other in-place tensor ops), the hyper-parameters that are baked into a
    plan (LIF tau / threshold / reset value, BatchNorm eps) and the device.  NOT covered: in-place edits through
    ``.data`` (``p.data.copy_()``, EMA swaps), which do not bump ``_version`` -- call ``invalidate_plans()`` after
    those."""
    hyper = []
    for m in module.modules():
        if hasattr(m, "v_threshold") and hasattr(m, "tau"):
            hyper.append((float(m.tau), float(m.v_threshold), None if m.v_reset is None else float(m.v_reset),
                          bool(getattr(m, "decay_input", True))))
        elif isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            hyper.append((float(m.eps),))
    dev = _lib.first_cuda_device(module)
    return (tuple(p._version for p in module.parameters()), tuple(b._version for b in module.buffers()),
            tuple(hyper), dev)


class PlanCacheMixin:
    """For nn.Modules that cache fused plans in ``self._plans``: explicit invalidation, and plans (ctypes function
    pointers, CUDA streams and graphs, device buffers) are kept out of pickling / copy.deepcopy / torch.save(module)."""

    def invalidate_plans(self) -> None:
        self._plans = {}

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_plans"] = {}
        return state


def _coef_array(coef: torch.Tensor, T: int):
    """The T memout coefficients as a host float array.  A coefficient buffer bound to another T (SURVEY.md finding 1:
    checkpoints carry ``memout.coef`` of shape (16,1,1,1,1)) raises the broadcasting RuntimeError the reference raises
    at R/snn_model/snn_layers.py:38 -- the fused plans must not silently truncate or zero-pad it."""
    c = coef.detach().reshape(-1).cpu().tolist()
    if len(c) != T:
        raise RuntimeError(f"The size of tensor a ({T}) must match the size of tensor b ({len(c)}) at "
                           "non-singleton dimension 0")
    return (ctypes.c_float * T)(*[float(v) for v in c])


def stf_empty(T: int, B: int, C: int, H: int, W: int, device) -> torch.Tensor:
    """Zero-filled STF buffer (pad/guard rows must stay zero; kernels only write valid rows)."""
    n = lib().sd_stf_bytes(T, B, C, H, W) // 2
    return torch.zeros(n, dtype=torch.float16, device=device)


@_lib.on_device_of
def stf_from_nchw(x: torch.Tensor) -> torch.Tensor:
    T, B, C, H, W = x.shape
    out = torch.empty(lib().sd_stf_bytes(T, B, C, H, W) // 2, dtype=torch.float16, device=x.device)
    check(lib().sd_stf_from_nchw(ptr(x.contiguous().float()), ptr(out), T, B, C, H, W, stream_ptr()))
    return out


@_lib.on_device_of
def stf8_from_nchw(x: torch.Tensor) -> torch.Tensor:
    """fp32 {0,1} spikes [T,B,C,H,W] -> STF8 (u8 planes of s and 128*s; the operand format of the kind::i8 layers)."""
    T, B, C, H, W = x.shape
    out = torch.zeros(lib().sd_stf_bytes(T, B, C, H, W), dtype=torch.uint8, device=x.device)   # guard / pad rows zero
    check(lib().sd_stf8_from_nchw(ptr(x.contiguous().float()), ptr(out), T, B, C, H, W, stream_ptr()))
    return out


@_lib.on_device_of
def stf8_to_nchw(stf8: torch.Tensor, T: int, B: int, C: int, H: int, W: int) -> torch.Tensor:
    out = torch.empty((T, B, C, H, W), dtype=torch.float32, device=stf8.device)
    check(lib().sd_stf8_to_nchw(ptr(stf8), ptr(out), T, B, C, H, W, stream_ptr()))
    return out


@_lib.on_device_of
def stf_to_nchw(stf: torch.Tensor, T: int, B: int, C: int, H: int, W: int) -> torch.Tensor:
    out = torch.empty((T, B, C, H, W), dtype=torch.float32, device=stf.device)
    check(lib().sd_stf_to_nchw(ptr(stf), ptr(out), T, B, C, H, W, stream_ptr()))
    return out


def fold_bn(conv_bias: Optional[torch.Tensor], c_out: int, bn, device):
    """(scale, shift) with BN(conv_nobias(x) + b) = conv_nobias(x) * scale + shift, evaluated in float64.

    SpikingJelly's own fold helper asserts ``conv.bias is None`` (SJ/activation_based/functional.py:731) while
    every reference conv has a bias, so the fold is re-derived:
    scale = gamma / sqrt(running_var + eps);  shift = (b - running_mean) * scale + beta.
    """
    b = torch.zeros(c_out, dtype=torch.float64, device=device) if conv_bias is None else conv_bias.detach().double()
    if bn is None:
        return torch.ones(c_out, dtype=torch.float32, device=device), b.float().contiguous()
    inv = torch.rsqrt(bn.running_var.detach().double() + bn.eps)
    gamma = bn.weight.detach().double() if bn.weight is not None else torch.ones_like(inv)
    beta = bn.bias.detach().double() if bn.bias is not None else torch.zeros_like(inv)
    scale = gamma * inv
    shift = (b - bn.running_mean.detach().double()) * scale + beta
    return scale.float().contiguous(), shift.float().contiguous()


class FusedLayer:
    """One conv[-BN][-LIF] stage bound to a fixed shape.  ``impl``: 'simt', 'tc' or 'auto'."""

    def __init__(self, conv, bn, lif, *, T: int, B: int, H_in: int, W_in: int, in_kind: int, out_kind: int,
                 impl: str = "auto", nsplit: int = 2, in_T: Optional[int] = None, C_in0: Optional[int] = None,
                 memout_coef: Optional[torch.Tensor] = None, share: Optional["FusedLayer"] = None,
                 concurrent: int = 1):
        L = lib()
        w = conv.weight.detach()
        _require_cuda(w, "layer weights")
        transposed = bool(getattr(conv, "transposed", isinstance(conv, torch.nn.ConvTranspose2d)))
        kh, kw = conv.kernel_size
        stride, pad = conv.stride[0], conv.padding[0]
        if conv.stride[0] != conv.stride[1] or conv.padding[0] != conv.padding[1]:
            raise ValueError("only square stride/padding are supported")
        if conv.dilation != (1, 1) or conv.groups != 1:
            raise ValueError("dilation/groups other than 1 are not supported")
        if transposed:
            C_in, C_out = w.shape[0], w.shape[1]
            op_h, op_w = conv.output_padding[0], conv.output_padding[1]
            H_out = (H_in - 1) * stride - 2 * pad + kh + op_h
            W_out = (W_in - 1) * stride - 2 * pad + kw + op_w
        else:
            C_out, C_in = w.shape[0], w.shape[1]
            H_out = (H_in + 2 * pad - kh) // stride + 1
            W_out = (W_in + 2 * pad - kw) // stride + 1
        d = ConvDesc()
        d.T, d.B, d.C_in, d.H_in, d.W_in = T, B, C_in, H_in, W_in
        d.C_out, d.H_out, d.W_out = C_out, H_out, W_out
        d.kh, d.kw, d.stride, d.pad, d.transposed = kh, kw, stride, pad, int(transposed)
        d.in_kind, d.out_kind = in_kind, out_kind
        d.in_T = (T if in_T is None else in_T) if in_kind == _lib.IN_STF else T
        d.C_in0 = C_in if C_in0 is None else C_in0
        if lif is not None:
            d.tau, d.v_threshold = float(lif.tau), float(lif.v_threshold)
            d.hard_reset = int(lif.v_reset is not None)
            d.v_reset = float(lif.v_reset) if lif.v_reset is not None else 0.0
            if not getattr(lif, "decay_input", True):
                raise ValueError("fused layers implement decay_input=True (the reference's setting)")
        else:
            d.tau, d.v_threshold, d.v_reset, d.hard_reset = 2.0, 1.0, 0.0, 1
        d.nsplit = nsplit
        d.concurrent = int(concurrent)
        if nsplit == 3 and impl != "simt":
            # kind::i8 layer: u8 spikes in (STF8); u8 spikes out, or -- the training branch's un-fused convolution -- the
            # fp32 currents in the planar row geometry (OUT_CURRENT_SEQ)
            if in_kind != _lib.IN_STF or out_kind not in (_lib.OUT_LIF, _lib.OUT_CURRENT_SEQ):
                raise ValueError("nsplit=3 (int8 weight digits) is a spike -> LIF (or spike -> current) layer")
            d.in_kind = _lib.IN_STF8
            d.out_kind = _lib.OUT_LIF8 if out_kind == _lib.OUT_LIF else _lib.OUT_CURRENT_SEQ
        if impl == "simt":
            d.nsplit = 2                      # unused by the CUDA-core kernels
        self.desc = d
        self.algorithmic_flops = None   # set when the layer runs a mathematically equal but larger problem
        self.device = w.device
        self.T, self.B, self.C_in, self.C_out = T, B, C_in, C_out
        self.H_in, self.W_in, self.H_out, self.W_out = H_in, W_in, H_out, W_out
        if impl == "auto":
            impl = "tc" if L.sd_conv_tc_supported(ctypes.byref(d)) else "simt"
        if impl == "tc" and not L.sd_conv_tc_supported(ctypes.byref(d)):
            raise ValueError("layer is not supported by the tcgen05 kernel: " + L.sd_last_error().decode())
        self.impl = impl
        self._fn = L.sd_conv_lif_tc if impl == "tc" else L.sd_conv_lif_simt
        self._coef = None
        ws = L.sd_conv_workspace_bytes(ctypes.byref(d)) if impl == "tc" else 0
        self._ws = torch.empty(ws, dtype=torch.uint8, device=self.device) if ws else None
        self.layout = L.sd_conv_weight_layout_tc(ctypes.byref(d)) if impl == "tc" else 0
        if (share is not None and share.impl == impl and share.desc.nsplit == d.nsplit and share.T == T
                and share.layout == self.layout):
            # sub-batch plans reuse the packed weights when their tile configuration (the layout key) is the same
            self.wpack, self.scale, self.shift, self._coef = share.wpack, share.scale, share.shift, share._coef
            return
        scale, shift = fold_bn(conv.bias, C_out, bn, self.device)
        wsrc = w.float().contiguous()
        if impl == "tc":
            nbytes = L.sd_conv_weight_bytes_tc(ctypes.byref(d))
            self.wpack = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            chan = torch.empty(C_out, dtype=torch.float32, device=self.device)
            check(L.sd_conv_pack_weights_tc(ctypes.byref(d), ptr(wsrc), ptr(self.wpack), ptr(chan), stream_ptr()))
            scale = (scale.double() * chan.double()).float().contiguous()  # exact: chan is a power of two
        else:
            nbytes = L.sd_conv_weight_bytes_simt(ctypes.byref(d))
            self.wpack = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            check(L.sd_conv_pack_weights_simt(ctypes.byref(d), ptr(wsrc), ptr(self.wpack), stream_ptr()))
        self.scale, self.shift = scale, shift
        if out_kind == _lib.OUT_MEMOUT_TANH:
            if memout_coef is None:
                raise ValueError("memout_coef required for the memout+tanh tail")
            self._coef = _coef_array(memout_coef, T)

    def flops(self) -> int:
        """Dense algorithmic FLOPs (2*MAC) of one call, counted as SURVEY.md section 8(d) does."""
        if self.algorithmic_flops is not None:
            return self.algorithmic_flops
        d = self.desc
        if d.transposed:
            mac = d.C_in * d.C_out * d.kh * d.kw * d.H_in * d.W_in
        else:
            mac = d.C_in * d.C_out * d.kh * d.kw * d.H_out * d.W_out
        return 2 * mac * d.B * d.T

    def executed_tensor_ops(self) -> int:
        """Tensor-pipe operations the layer really issues (2 x MAC of the problem as run x weight terms): the upsampled /
        stride-1 restatements run 4x the algorithm's MACs, the T-summed read-out 1/T of them, exact weights 2 fp16
        terms (or 3 int8 digits) per algorithmic MAC.  0 for the CUDA-core kernels."""
        if self.impl != "tc":
            return 0
        d = self.desc
        mac = d.C_in * d.C_out * d.kh * d.kw * d.H_out * d.W_out * d.B * (d.in_T if d.in_kind == _lib.IN_STF else d.T)
        return 2 * mac * d.nsplit

    def mma_kind(self) -> str:
        if self.impl != "tc":
            return "cuda cores (fp32)"
        return {1: "tcgen05 kind::f16, 1 fp16 weight term", 2: "tcgen05 kind::f16, 2 exact fp16 weight terms (2 passes)",
                3: "tcgen05 kind::i8, 3 int8 weight digits at twice the rate (1.5 pass equivalents), exact int32 accumulation"}[self.desc.nsplit]

    def alloc_out(self) -> torch.Tensor:
        d = self.desc
        if d.out_kind == _lib.OUT_LIF8:      # same byte count as the fp16 format; guard / pad rows must be zero
            return torch.zeros(lib().sd_stf_bytes(d.T, d.B, d.C_out, d.H_out, d.W_out), dtype=torch.uint8, device=self.device)
        if d.out_kind == _lib.OUT_LIF:
            return stf_empty(d.T, d.B, d.C_out, d.H_out, d.W_out, self.device)
        if d.out_kind == _lib.OUT_REAL_SEQ:
            return torch.empty((d.T, d.B, d.C_out, d.H_out, d.W_out), dtype=torch.float32, device=self.device)
        if d.out_kind == _lib.OUT_CURRENT_SEQ:   # [T][C_out/8][R_alloc][8] fp32: twice the bytes of the fp16 STF tensor
            return torch.empty(lib().sd_stf_bytes(d.T, d.B, d.C_out, d.H_out, d.W_out) // 2, dtype=torch.float32,
                               device=self.device)
        if d.out_kind == _lib.OUT_MEMOUT_TANH:
            return torch.empty((d.B, d.C_out, d.H_out, d.W_out), dtype=torch.float32, device=self.device)
        return torch.empty((d.B, d.H_out, d.W_out, d.C_out), dtype=torch.float32, device=self.device)

    def alloc_sum(self) -> torch.Tensor:
        d = self.desc
        return stf_empty(1, d.B, d.C_out, d.H_out, d.W_out, self.device)

    def alloc_state(self) -> torch.Tensor:
        """LIF membrane state in the planar layout [C_out/8][R_alloc][8] fp32, initialised to v_reset."""
        d = self.desc
        n = lib().sd_stf_bytes(1, d.B, d.C_out, d.H_out, d.W_out) // 2
        return torch.full((n,), d.v_reset if d.hard_reset else 0.0, dtype=torch.float32, device=self.device)

    def run(self, x: torch.Tensor, out: torch.Tensor, x2: Optional[torch.Tensor] = None,
            out_sum: Optional[torch.Tensor] = None, v: Optional[torch.Tensor] = None, in_scalar: float = 0.0) -> torch.Tensor:
        a = ConvArgs()
        a.in_scalar = float(in_scalar)
        a.in_, a.in2, a.weights = ptr(x), ptr(x2), ptr(self.wpack)
        a.scale, a.shift, a.v = ptr(self.scale), ptr(self.shift), ptr(v)
        a.out, a.out_sum = ptr(out), ptr(out_sum)
        a.memout_coef_host = ctypes.cast(self._coef, ctypes.c_void_p) if self._coef is not None else None
        a.workspace = ptr(self._ws)
        check(self._fn(ctypes.byref(self.desc), ctypes.byref(a), stream_ptr()))
        return out


# --------------------------------------------------------------------------------------------------
class DenoiserPlan:
    """DummyModel.forward for a fixed (T, b, h, w): conv1 (real, constant over T) on CUDA cores, conv2..conv5 as
    fused tcgen05 conv+BN+LIF, conv6 on the T-summed spikes of cat(x5, x1) (linear read-out, mean over T)."""

    def __init__(self, model, T: int, b: int, h: int, w: int, nsplit: int = 2, impl: str = "auto",
                 weights_from: Optional["DenoiserPlan"] = None, concurrent: int = 1):
        dev = model.conv1[0].weight.device
        _require_cuda(model.conv1[0].weight, "DummyModel parameters")
        self.T, self.b, self.h, self.w, self.K = T, b, h, w, model.num_embeddings
        self.device = dev
        wf = weights_from
        mk = lambda seq, sh, **kw: FusedLayer(seq[0], seq[1] if len(seq) > 1 else None, seq[2] if len(seq) > 2 else None,
                                              T=T, B=b, H_in=h, W_in=w, nsplit=nsplit, share=sh,
                                              concurrent=concurrent, **kw)
        # nsplit = 3: conv2..conv5 on the kind::i8 tensor-core path (u8 spikes between the layers, three int8 weight
        # digits, exact int32 accumulation: 3 MMAs per 32 input channels where two fp16 terms need 4).  It needs an even
        # T; the read-out layer consumes T-summed counts (up to T per element) and stays on the fp16 path.
        self.i8 = nsplit == 3 and T % 2 == 0 and impl != "simt" and self._int8_layers_supported(model, T, b, h, w)
        ns = 3 if self.i8 else (2 if nsplit == 3 else nsplit)
        self.l1 = mk(model.conv1, wf and wf.l1, in_kind=_lib.IN_REAL_CONST,
                     out_kind=_lib.OUT_LIF8 if self.i8 else _lib.OUT_LIF, impl="simt")
        # the sampler's entry: conv1 reads the int64 token grid and the scalar diffusion time directly
        # (cat(x, t * ones) -> repeat(T) of vq_diffusion.py:195-198 is never materialised); same packed weights
        self.l1t = mk(model.conv1, self.l1, in_kind=_lib.IN_TOKENS,
                      out_kind=_lib.OUT_LIF8 if self.i8 else _lib.OUT_LIF, impl="simt")
        mk_s = lambda seq, sh: FusedLayer(seq[0], seq[1], seq[2], T=T, B=b, H_in=h, W_in=w, nsplit=ns, share=sh,
                                          concurrent=concurrent, in_kind=_lib.IN_STF, out_kind=_lib.OUT_LIF, impl=impl)
        self.l2 = mk_s(model.conv2, wf and wf.l2)
        self.l3 = mk_s(model.conv3, wf and wf.l3)
        self.l4 = mk_s(model.conv4, wf and wf.l4)
        self.l5 = mk_s(model.conv5, wf and wf.l5)
        self.l6 = FusedLayer(model.conv6[0], None, None, T=T, B=b, H_in=h, W_in=w, nsplit=2 if ns == 3 else ns,
                             share=wf and wf.l6, concurrent=concurrent, in_kind=_lib.IN_STF, out_kind=_lib.OUT_MEAN_T,
                             impl=impl, in_T=1, C_in0=256)
        self.layers = [self.l1, self.l2, self.l3, self.l4, self.l5, self.l6]
        self.xin = torch.empty((b, 2, h, w), dtype=torch.float32, device=dev)
        self.x1, self.x1s = self.l1.alloc_out(), self.l1.alloc_sum()
        self.x2, self.x3, self.x4 = self.l2.alloc_out(), self.l3.alloc_out(), self.l4.alloc_out()
        self.x5, self.x5s = self.l5.alloc_out(), self.l5.alloc_sum()
        self.logits = self.l6.alloc_out()  # [b, h, w, K] channels last

    @staticmethod
    def _int8_layers_supported(model, T, b, h, w) -> bool:
        """Whether sd_conv_lif_tc takes conv2..conv5 of this shape as kind::i8 layers (channel multiples, grid width); if not,
        the plan uses two fp16 terms (or the CUDA-core kernels) exactly as with nsplit = 2."""
        for seq in (model.conv2, model.conv3, model.conv4, model.conv5):
            conv = seq[0]
            d = ConvDesc()
            d.T, d.B, d.C_in, d.H_in, d.W_in = T, b, conv.in_channels, h, w
            d.C_out, d.H_out, d.W_out = conv.out_channels, h, w
            d.kh, d.kw, d.stride, d.pad, d.transposed = conv.kernel_size[0], conv.kernel_size[1], conv.stride[0], conv.padding[0], 0
            d.in_kind, d.out_kind, d.in_T, d.C_in0 = _lib.IN_STF8, _lib.OUT_LIF8, T, conv.in_channels
            d.tau, d.v_threshold, d.v_reset, d.hard_reset, d.nsplit, d.concurrent = 2.0, 1.0, 0.0, 1, 3, 1
            if not lib().sd_conv_tc_supported(ctypes.byref(d)):
                return False
        return True

    def flops(self) -> int:
        return sum(l.flops() for l in self.layers)

    def spikes_nchw(self, buf: torch.Tensor, lyr: "FusedLayer") -> torch.Tensor:
        """The spikes a fused layer left in its inter-layer buffer, as the reference's fp32 [T, b, C, h, w] tensor
        (diagnostics / parity tests; the product path never converts)."""
        if lyr.desc.out_kind == _lib.OUT_LIF8:
            return stf8_to_nchw(buf, lyr.T, lyr.B, lyr.C_out, lyr.H_out, lyr.W_out)
        return stf_to_nchw(buf, lyr.T, lyr.B, lyr.C_out, lyr.H_out, lyr.W_out)

    def run_from_input(self, tokens: Optional[torch.Tensor] = None, t: float = 0.0) -> torch.Tensor:
        if tokens is None:
            self.l1.run(self.xin, self.x1, out_sum=self.x1s)
        else:
            self.l1t.run(tokens, self.x1, out_sum=self.x1s, in_scalar=t)
        self.l2.run(self.x1, self.x2)
        self.l3.run(self.x2, self.x3)
        self.l4.run(self.x3, self.x4)
        self.l5.run(self.x4, self.x5, out_sum=self.x5s)
        self.l6.run(self.x5s, self.logits, x2=self.x1s)
        return self.logits

    def run_tokens(self, x_t: torch.Tensor, t: int) -> torch.Tensor:
        """x_t int64 [b*h*w] token ids, scalar diffusion time t (the sampler uses one t for the whole batch)."""
        return self.run_from_input(x_t, float(int(t)))

    def run(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """General entry: x float [b,1,h,w], t long [b] (per-sample times, as DummyModel.forward allows)."""
        self.xin[:, 0:1].copy_(x)
        self.xin[:, 1:2].copy_(t.to(self.xin.dtype).reshape(-1, 1, 1, 1).expand(-1, 1, self.h, self.w))
        return self.run_from_input()


def currents_to_nchw(cur: torch.Tensor, T: int, B: int, C: int, H: int, W: int) -> torch.Tensor:
    """OUT_CURRENT_SEQ buffer ([T][C/8][R_alloc][8] fp32, rows = guard + b*H*W + y*W + x) -> [T, B, C, H, W] fp32.
    One strided copy; used by the training branch only (include/sd_b200.h: STF geometry)."""
    guard = (W + 1 + 7) // 8 * 8
    rows = cur.numel() // (T * (C // 8) * 8)
    v = cur.view(T, C // 8, rows, 8)[:, :, guard:guard + B * H * W]
    return v.reshape(T, C // 8, B, H, W, 8).permute(0, 2, 1, 5, 3, 4).reshape(T, B, C, H, W).contiguous()


def cluster_count() -> int:
    """2-SM clusters of the current device (sm_count / 2: 74 on a B200); 74 when no device is present (host-side
    planning tests run without a GPU)."""
    sms = ctypes.c_int(0)
    try:
        if lib().sd_device_info(ctypes.byref(sms), None, None, None) == 0 and sms.value > 0:
            return sms.value // 2
    except Exception:  # noqa: BLE001  (library not built: the caller fails loudly at its first kernel call)
        pass
    return 74


def plan_sub_batches(b: int, rows_per_image: int, n_streams: Optional[int] = None) -> list:
    """Sizes of the concurrent sub-batches of a shard of ``b`` images (measured on B200, profiles/r01_experiments.md).

    The tcgen05 layers work on pairs of 128-row tiles (256 rows of the dense b*h*w pixel grid), one pair per 2-SM
    cluster, sm_count / 2 = 74 clusters on a B200:
     * a shard of at most one wave of pairs runs best as up to 5 concurrent sub-batches of >= 5 pairs (their layers
       pack the SMs like small items pack a bin); larger shards as 2 (fewer, longer launches; the GPU is power-capped
       there); below 64 images sub-batches would only add launches;
     * sub-batch sizes are cut so that their row count ends just below a multiple of 256: no half-empty pair.
    ``n_streams`` (or env SD_SAMPLER_STREAMS) overrides the count; the alignment rule still applies."""
    total_pairs = -(-b * rows_per_image // 256)
    if n_streams is None:
        env = os.environ.get("SD_SAMPLER_STREAMS")
        n_streams = int(env) if env else (min(5, max(1, total_pairs // 5)) if total_pairs <= cluster_count() else 2)
    if b < 64:
        n_streams = 1
    n_streams = max(1, min(n_streams, b))
    if n_streams == 1:
        return [b]
    pairs_per_sub = -(-total_pairs // n_streams)
    per = max(1, (pairs_per_sub * 256) // rows_per_image)
    sizes = []
    lo = 0
    while lo < b:
        sizes.append(min(per, b - lo))
        lo += sizes[-1]
    return sizes


class SamplerPlan:
    """AbsorbingDiffusion.sample for a fixed batch shard.

    Philox stream: step i draws its uniforms at generator offset ``offset0 + i*(inc_u + inc_e)`` and its
    exponentials at ``+ inc_u``, exactly what two consecutive torch CUDA calls (rand_like, exponential_) consume for
    the GLOBAL batch; a shard [token_base, token_base + n) evaluates only its own elements of that stream.

    Sub-batches and streams: images are independent, so the shard is cut into ``n_streams`` sub-batches whose whole
    reverse-diffusion loops run on separate CUDA streams.  Each fused layer is a persistent kernel with one CTA per
    SM; with a single stream the last, partially filled wave of every layer idles SMs (cfg2: 49 tile pairs x 2 N
    tiles on 74 clusters = 1.32 waves).  With several streams the tail of one sub-batch's layer overlaps the other
    sub-batches' layers, and the small kernels (input, conv1, sampling step) hide behind the large ones.
    """

    def __init__(self, model, T: int, b: int, h: int, w: int, mask_id: int, n_global: Optional[int] = None,
                 shard_base: int = 0, n_streams: Optional[int] = None, nsplit: int = 2):
        sizes = plan_sub_batches(b, h * w, n_streams)
        n_streams = len(sizes)
        self.mask_id = int(mask_id)
        self.b, self.h, self.w, self.K = b, h, w, model.num_embeddings
        hw = h * w
        self.n_tokens = b * hw
        self.n_tokens_global = (n_global if n_global is not None else b) * hw
        self.token_base = shard_base * hw
        dev = model.conv1[0].weight.device
        self.device = dev
        self.x_t = torch.empty(self.n_tokens, dtype=torch.int64, device=dev)
        self.unmasked = torch.empty(self.n_tokens, dtype=torch.uint8, device=dev)
        self.subs = []
        lo = 0
        for bi in sizes:
            # each sub-batch has its own activation buffers (they run concurrently) but shares the packed weights
            dp = DenoiserPlan(model, T, bi, h, w, nsplit=nsplit, weights_from=self.subs[0][0] if self.subs else None,
                              concurrent=n_streams)
            self.subs.append((dp, lo, bi))
            lo += bi
        self.dp = self.subs[0][0]
        self.streams = [torch.cuda.Stream(device=dev) for _ in self.subs] if len(self.subs) > 1 else [None]
        inc = ctypes.c_uint64()
        check(lib().sd_philox_offset_increment(self.n_tokens_global, ctypes.byref(inc)))
        self.inc_u = inc.value
        check(lib().sd_philox_offset_increment(self.n_tokens_global * self.K, ctypes.byref(inc)))
        self.inc_e = inc.value
        self.kernel_launches_per_step = 7 * len(self.subs)     # conv1 (token input) .. conv6, sampling step

    def flops_per_image(self, sample_steps: int) -> int:
        return sum(dp.flops() for dp, _, _ in self.subs) * sample_steps // self.b

    def offset_advance(self, sample_steps: int) -> int:
        return sample_steps * (self.inc_u + self.inc_e)

    def _step(self, dp, lo, bi, t, temp, seed, off, rng_dev=None, hist_row=None):
        hw = self.h * self.w
        xs = self.x_t[lo * hw:(lo + bi) * hw]
        us = self.unmasked[lo * hw:(lo + bi) * hw]
        logits = dp.run_tokens(xs, t)
        if rng_dev is None:
            check(lib().sd_sample_step(ptr(logits), ptr(xs), ptr(us), None, bi * hw, self.K, t, float(temp), int(seed),
                                       off, off + self.inc_u, self.token_base + lo * hw, self.n_tokens_global,
                                       stream_ptr()))
        else:  # `off` is relative to the base offset stored in rng_dev
            check(lib().sd_sample_step_dev(ptr(logits), ptr(xs), ptr(us), None, bi * hw, self.K, t, float(temp),
                                           ptr(rng_dev), off, off + self.inc_u, self.token_base + lo * hw,
                                           self.n_tokens_global, stream_ptr()))
        if hist_row is not None:   # parity diagnostics: the token grid after this step (tests/test_gpu_sample_parity.py)
            hist_row[lo * hw:(lo + bi) * hw].copy_(xs)

    def _enqueue(self, temp, sample_steps, seed, offset0, fill_x, fill_u, rng_dev=None, history=None):
        """Enqueue the whole reverse-diffusion loop on the current stream (+ the sub-batch streams)."""
        if fill_x:
            self.x_t.fill_(self.mask_id)
        if fill_u:
            self.unmasked.zero_()
        multi = len(self.subs) > 1
        if multi:
            cur = torch.cuda.current_stream()
            start = torch.cuda.Event()
            start.record(cur)
            for s in self.streams:
                s.wait_event(start)
        off = int(offset0)
        for t in range(sample_steps, 0, -1):
            row = None if history is None else history[sample_steps - t]
            for (dp, lo, bi), s in zip(self.subs, self.streams):
                if multi:
                    with torch.cuda.stream(s):
                        self._step(dp, lo, bi, t, temp, seed, off, rng_dev, row)
                else:
                    self._step(dp, lo, bi, t, temp, seed, off, rng_dev, row)
            off += self.inc_u + self.inc_e
        if multi:
            for s in self.streams:
                done = torch.cuda.Event()
                done.record(s)
                cur.wait_event(done)

    def sample(self, temp: float, sample_steps: int, seed: int, offset0: int = 0,
               x_init: Optional[torch.Tensor] = None, unmasked_init: Optional[torch.Tensor] = None,
               use_graph: Optional[bool] = None, history: Optional[torch.Tensor] = None,
               extra_base: int = 0) -> torch.Tensor:
        """x_init / unmasked_init (host or device, [b,1,h,w] or flat): start from a partially unmasked grid
        instead of the all-mask grid of vq_diffusion.py:106-107; copied on the current stream (pinned host -> device).

        extra_base: images added to the plan's shard_base for this call only (a large batch generated as consecutive
        chunks by ONE plan and one captured graph; the value lives in device memory next to the RNG state).

        use_graph (default: env SD_SAMPLER_GRAPH, on): the loop (h*w steps x 8 kernels x sub-batches) is captured once
        per (temp, steps) into a CUDA graph and replayed; the Philox (seed, offset) pair lives in device memory
        (sd_sample_step_dev) so every replay draws a fresh, torch-identical stream."""
        import numpy as np
        if use_graph is None:
            use_graph = os.environ.get("SD_SAMPLER_GRAPH", "1") != "0"
        if x_init is not None:
            self.x_t.copy_(x_init.reshape(-1), non_blocking=True)
        if unmasked_init is not None:
            self.unmasked.copy_(unmasked_init.reshape(-1), non_blocking=True)
        if history is not None:
            # int64 [sample_steps, b*h*w] device buffer receiving the token grid after every step (diagnostics; eager)
            if history.shape != (sample_steps, self.n_tokens) or history.dtype != torch.int64 or not history.is_cuda:
                raise ValueError(f"history must be a CUDA int64 tensor of shape ({sample_steps}, {self.n_tokens})")
            use_graph = False
        extra_tokens = int(extra_base) * self.h * self.w
        if self.token_base + extra_tokens + self.n_tokens > self.n_tokens_global:
            raise ValueError("extra_base moves the shard past the end of the global batch")
        if not use_graph:
            saved = self.token_base
            self.token_base += extra_tokens
            try:
                self._enqueue(temp, sample_steps, seed, offset0, x_init is None, unmasked_init is None, history=history)
            finally:
                self.token_base = saved
            return self.x_t.view(self.b, 1, self.h, self.w)
        if not hasattr(self, "rng_dev"):
            self.rng_dev = torch.zeros(3, dtype=torch.int64, device=self.x_t.device)
            self._graphs = {}
        state = np.array([int(seed) & 0xFFFFFFFFFFFFFFFF, int(offset0), extra_tokens], dtype=np.uint64).view(np.int64)
        self.rng_dev.copy_(torch.from_numpy(state))
        key = (float(temp), int(sample_steps), x_init is None, unmasked_init is None)
        g = self._graphs.get(key)
        if g is None:
            # one eager step first: lazy one-time initialisation (function attributes) must not happen under capture
            xt_save, um_save = self.x_t.clone(), self.unmasked.clone()
            self._enqueue(temp, 1, seed, 0, False, False, self.rng_dev)
            self.x_t.copy_(xt_save)
            self.unmasked.copy_(um_save)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue(temp, sample_steps, 0, 0, x_init is None, unmasked_init is None, self.rng_dev)
            self._graphs[key] = g
        g.replay()
        return self.x_t.view(self.b, 1, self.h, self.w)


# --------------------------------------------------------------------------------------------------
class _UpsampledConvT:
    """ConvTranspose2d(k=3, s=2, p=1, output_padding=1) restated as the stride-1 3x3 convolution (pad 1) it equals on the
    zero-inserted 2x upsampled input: out[o] = sum_i in[i] w[o + 1 - 2i]  ==  sum_k' u[o - 1 + k'] w[2 - k'] with
    u[2i] = in[i].  Carries the attributes FusedLayer reads from a conv module."""

    def __init__(self, convt):
        # [C_in, C_out, 3, 3] -> [C_out, C_in, 3, 3], taps flipped
        self.weight = convt.weight.detach().transpose(0, 1).flip(2, 3).contiguous()
        self.bias = convt.bias
        self.kernel_size, self.stride, self.padding, self.dilation, self.groups = (3, 3), (1, 1), (1, 1), (1, 1), 1
        self.transposed = False
        self.in_channels, self.out_channels = convt.in_channels, convt.out_channels

    @staticmethod
    def eligible(convt) -> bool:
        return (bool(getattr(convt, "transposed", isinstance(convt, torch.nn.ConvTranspose2d)))
                and tuple(convt.kernel_size) == (3, 3) and tuple(convt.stride) == (2, 2)
                and tuple(convt.padding) == (1, 1) and tuple(convt.output_padding) == (1, 1)
                and convt.in_channels % 16 == 0 and convt.out_channels % 16 == 0)


class _Stride1Conv:
    """Conv2d(k=3, s=2, p=1) restated as the stride-1 convolution whose even output positions it is."""

    def __init__(self, conv):
        self.weight, self.bias = conv.weight, conv.bias
        self.kernel_size, self.stride, self.padding, self.dilation, self.groups = (3, 3), (1, 1), (1, 1), (1, 1), 1
        self.transposed = False
        self.in_channels, self.out_channels = conv.in_channels, conv.out_channels

    @staticmethod
    def eligible(conv) -> bool:
        return (not bool(getattr(conv, "transposed", isinstance(conv, torch.nn.ConvTranspose2d)))
                and tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (2, 2)
                and tuple(conv.padding) == (1, 1) and tuple(conv.dilation) == (1, 1) and conv.groups == 1
                and conv.in_channels % 16 == 0 and conv.out_channels % 16 == 0)


class VQVAEPlan:
    """Encoder -> quantiser -> spike generator -> decoder of SNN_VQVAE for a fixed (T, B, H, W)."""

    def __init__(self, model, T: int, B: int, H: int, W: int, nsplit: int = 2):
        enc, dec, vq = model.encoder.snn_convs, model.decoder.snn_convs, model.vq_layer
        dev = enc[0].weight.device
        _require_cuda(enc[0].weight, "SNN_VQVAE parameters")
        self.T, self.B, self.H, self.W, self.device = T, B, H, W, dev
        self.D, self.K = vq.embedding_dim, vq.num_embeddings
        self.in_dim = enc[0].in_channels
        mk = lambda conv, bn, lif, Hi, Wi, **kw: FusedLayer(conv, bn, lif, T=T, B=B, H_in=Hi, W_in=Wi, nsplit=nsplit, **kw)
        self.e1 = mk(enc[0], enc[1], enc[2], H, W, in_kind=_lib.IN_REAL_SEQ, out_kind=_lib.OUT_LIF, impl="simt")
        self.e1c = mk(enc[0], enc[1], enc[2], H, W, in_kind=_lib.IN_REAL_CONST, out_kind=_lib.OUT_LIF, impl="simt")
        # enc.conv2 (stride 2, spike input): on the tcgen05 kernel at stride 1, even positions kept afterwards
        # (4x the MMAs of the algorithm, several times faster than CUDA cores); SD_ENCODER_TC=0 keeps the CUDA-core kernel
        self.tc_encoder = (os.environ.get("SD_ENCODER_TC", "1") != "0" and _Stride1Conv.eligible(enc[3])
                           and self.e1.W_out + 2 <= 64)
        if self.tc_encoder:
            self.e2 = mk(_Stride1Conv(enc[3]), enc[4], enc[5], self.e1.H_out, self.e1.W_out, in_kind=_lib.IN_STF,
                         out_kind=_lib.OUT_LIF)
            self.tc_encoder = self.e2.impl == "tc"
        if self.tc_encoder:
            self.e2.algorithmic_flops = self.e2.flops() // 4
            self.s2_full = self.e2.alloc_out()
            self.e2.H_out, self.e2.W_out = (self.e1.H_out + 1) // 2, (self.e1.W_out + 1) // 2
        else:
            self.e2 = mk(enc[3], enc[4], enc[5], self.e1.H_out, self.e1.W_out, in_kind=_lib.IN_STF, out_kind=_lib.OUT_LIF)
        self.e3 = mk(enc[6], enc[7], enc[8], self.e2.H_out, self.e2.W_out, in_kind=_lib.IN_STF, out_kind=_lib.OUT_LIF)
        self.h, self.w = self.e3.H_out, self.e3.W_out
        self.gen = mk(vq.poisson[0], vq.poisson[1], vq.poisson[2], self.h, self.w, in_kind=_lib.IN_REAL_CONST,
                      out_kind=_lib.OUT_LIF, impl="simt")
        # Decoder: the two stride-2 transposed convolutions run on the tcgen05 kernel as stride-1 convolutions of the
        # zero-inserted upsampled spikes (4x the MMAs of the algorithm, still several times faster than CUDA cores);
        # SD_DECODER_TC=0 keeps the CUDA-core kernels (used by the tests to cross-check the two paths).
        self.tc_decoder = (os.environ.get("SD_DECODER_TC", "1") != "0" and _UpsampledConvT.eligible(dec[0])
                           and _UpsampledConvT.eligible(dec[3]) and 4 * self.w + 2 <= 64)
        if self.tc_decoder:
            self.d1 = mk(_UpsampledConvT(dec[0]), dec[1], dec[2], 2 * self.h, 2 * self.w, in_kind=_lib.IN_STF,
                         out_kind=_lib.OUT_LIF)
            self.d2 = mk(_UpsampledConvT(dec[3]), dec[4], dec[5], 4 * self.h, 4 * self.w, in_kind=_lib.IN_STF,
                         out_kind=_lib.OUT_LIF)
            self.tc_decoder = self.d1.impl == "tc" and self.d2.impl == "tc"
        if self.tc_decoder:
            self.d1.algorithmic_flops = self.d1.flops() // 4
            self.d2.algorithmic_flops = self.d2.flops() // 4
            self.up0 = stf_empty(T, B, dec[0].in_channels, 2 * self.h, 2 * self.w, dev)
            self.up1 = stf_empty(T, B, dec[3].in_channels, 4 * self.h, 4 * self.w, dev)
        else:
            self.d1 = mk(dec[0], dec[1], dec[2], self.h, self.w, in_kind=_lib.IN_STF, out_kind=_lib.OUT_LIF)
            self.d2 = mk(dec[3], dec[4], dec[5], self.d1.H_out, self.d1.W_out, in_kind=_lib.IN_STF,
                         out_kind=_lib.OUT_LIF)
        self.d3 = mk(dec[6], None, None, self.d2.H_out, self.d2.W_out, in_kind=_lib.IN_STF,
                     out_kind=_lib.OUT_MEMOUT_TANH, memout_coef=model.memout.coef)
        self.s1, self.s3 = self.e1.alloc_out(), self.e3.alloc_out()
        self.s2 = stf_empty(T, B, self.e2.C_out, self.e2.H_out, self.e2.W_out, dev)
        self.z = torch.empty((B * self.h * self.w, self.D), dtype=torch.float32, device=dev)
        self.idx = torch.empty(B * self.h * self.w, dtype=torch.int64, device=dev)
        self.margin = torch.empty(B * self.h * self.w, dtype=torch.float32, device=dev)
        self.q = torch.empty((B, self.D, self.h, self.w), dtype=torch.float32, device=dev)
        self.sg, self.sd1, self.sd2 = self.gen.alloc_out(), self.d1.alloc_out(), self.d2.alloc_out()
        self.recon = self.d3.alloc_out()
        self._vq = vq
        self._states, self._captured = None, None
        self._coef_vq = _coef_array(vq.memout.coef, T)

    def _state(self, name: str, lyr: "FusedLayer") -> Optional[torch.Tensor]:
        """Fresh LIF state buffer (planar layout, reset value) for layer ``name`` while states are being captured."""
        if self._states is None:
            return None
        self._states[name] = (lyr.alloc_state(), lyr)
        return self._states[name][0]

    def encode(self, x: torch.Tensor, const_over_T: bool = False) -> torch.Tensor:
        """x: fp32 [T,B,C,H,W] (or [B,C,H,W] with const_over_T=True: the frame repeated T times, R/main.py:133)."""
        if const_over_T:
            self.e1c.run(x.contiguous(), self.s1, v=self._state("e1", self.e1c))
        else:
            self.e1.run(x.contiguous(), self.s1, v=self._state("e1", self.e1))
        if self.tc_encoder:
            self.e2.run(self.s1, self.s2_full, v=self._state("e2", self.e2))
            check(lib().sd_stf_subsample2x(ptr(self.s2_full), ptr(self.s2), self.T, self.B, self.e2.C_out,
                                           self.e1.H_out, self.e1.W_out, stream_ptr()))
        else:
            self.e2.run(self.s1, self.s2, v=self._state("e2", self.e2))
        self.e3.run(self.s2, self.s3, v=self._state("e3", self.e3))
        return self.s3

    def quantize_indices(self, spikes_stf: torch.Tensor) -> torch.Tensor:
        L, vq = lib(), self._vq
        check(L.sd_vq_feature(ptr(spikes_stf), ptr(vq.alpha.detach()), ctypes.cast(self._coef_vq, ctypes.c_void_p),
                              ptr(self.z), self.T, self.B, self.D, self.h, self.w, stream_ptr()))
        check(L.sd_vq_lookup(ptr(self.z), ptr(vq.embeddings.weight.detach()), ptr(self.idx), ptr(self.margin),
                             self.z.shape[0], self.D, self.K, stream_ptr()))
        return self.idx

    def generate(self, idx: torch.Tensor) -> torch.Tensor:
        """code indices [B*h*w] -> generator spikes (STF): quantize -> NCHW -> poisson (vae_model.py:50-57)."""
        L, vq = lib(), self._vq
        check(L.sd_vq_gather(ptr(idx), ptr(vq.embeddings.weight.detach()), ptr(self.q), self.B, self.D, self.h, self.w,
                             self.K, stream_ptr()))
        self.gen.run(self.q, self.sg, v=self._state("gen", self.gen))
        return self.sg

    def decode(self, e_stf: torch.Tensor) -> torch.Tensor:
        if self.tc_decoder:
            L, d1, d2 = lib(), self.d1.desc, self.d2.desc
            check(L.sd_stf_upsample2x(ptr(e_stf), ptr(self.up0), self.T, self.B, d1.C_in, self.h, self.w, stream_ptr()))
            self.d1.run(self.up0, self.sd1, v=self._state("d1", self.d1))
            check(L.sd_stf_upsample2x(ptr(self.sd1), ptr(self.up1), self.T, self.B, d2.C_in, 2 * self.h, 2 * self.w,
                                      stream_ptr()))
            self.d2.run(self.up1, self.sd2, v=self._state("d2", self.d2))
        else:
            self.d1.run(e_stf, self.sd1, v=self._state("d1", self.d1))
            self.d2.run(self.sd1, self.sd2, v=self._state("d2", self.d2))
        self.d3.run(self.sd2, self.recon)
        return self.recon

    def decode_indices(self, idx: torch.Tensor) -> torch.Tensor:
        """R/main.py:388-399: sampled indices -> tanh(memout(decoder(poisson(quantize(idx)))))."""
        return self.decode(self.generate(idx.reshape(-1)))

    def forward(self, x: torch.Tensor, const_over_T: bool = False, capture_states: bool = False):
        """Returns (generator spikes STF, reconstruction, code indices).  With ``capture_states`` the final membrane
        potential of every LIF layer is kept (``self.states()``) so that a caller following the reference's state
        protocol can hand it to the modules; otherwise the states are consumed inside the kernels."""
        self._states = {} if capture_states else None
        try:
            z = self.encode(x, const_over_T)
            idx = self.quantize_indices(z)
            e = self.generate(idx)
            rec = self.decode(e)
        finally:
            captured, self._states = self._states, None
        self._captured = captured
        return e, rec, idx

    def states(self):
        """{layer name: zero-argument function returning the final LIF state as fp32 [B, C, H, W]} of the last
        ``forward(capture_states=True)``.  The conversion from the planar kernel layout runs when the function is
        called; the tensor-core encoder route computes enc.conv2 at stride 1, so its state is the even positions."""
        out = {}
        for name, (buf, lyr) in (self._captured or {}).items():
            d = lyr.desc

            def make(buf=buf, d=d, sub=(name == "e2" and self.tc_encoder)):
                def fn():
                    v = torch.empty((d.B, d.C_out, d.H_out, d.W_out), dtype=torch.float32, device=buf.device)
                    check(lib().sd_state_convert(ptr(buf), ptr(v), d.B, d.C_out, d.H_out, d.W_out, 0, stream_ptr()))
                    return v[..., ::2, ::2].contiguous() if sub else v
                return fn
            out[name] = make()
        return out

    def flops(self) -> int:
        return sum(l.flops() for l in (self.e1, self.e2, self.e3, self.gen, self.d1, self.d2, self.d3))


@_lib.on_device_of
def to_uint8(pred: torch.Tensor) -> torch.Tensor:
    out = torch.empty(pred.shape, dtype=torch.uint8, device=pred.device)
    check(lib().sd_to_uint8(ptr(pred.contiguous()), ptr(out), pred.numel(), stream_ptr()))
    return out
