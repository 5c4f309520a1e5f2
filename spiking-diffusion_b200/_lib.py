"""ctypes binding of libsd_b200.so (the C-ABI declared in include/sd_b200.h) and its in-tree build.

PyTorch is plumbing here: it owns device memory and streams; every compute call goes through one C-ABI
symbol with raw ``tensor.data_ptr()`` pointers and the current CUDA stream.  There is no CPU fallback: if the
shared library is missing, or the process has no sm_100 device, calls raise.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libsd_b200.so")
SOURCES = ["lif.cu", "vq.cu", "conv_simt.cu", "conv_tc.cu", "sample.cu", "train.cu", "metrics.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--extended-lambda",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]

# Exported symbols of include/sd_b200.h (tests check that the library exports exactly these).
SYMBOLS = [
    "sd_last_error", "sd_version", "sd_device_info", "sd_stf_guard", "sd_stf_rows", "sd_stf_bytes",
    "sd_stf_from_nchw", "sd_stf_to_nchw", "sd_stf8_from_nchw", "sd_stf8_to_nchw", "sd_stf_upsample2x", "sd_stf_subsample2x", "sd_channel_affine", "sd_state_convert", "sd_lif_forward", "sd_lif_backward", "sd_memout", "sd_vq_feature", "sd_vq_lookup",
    "sd_vq_gather", "sd_conv_weight_bytes_simt", "sd_conv_weight_bytes_tc", "sd_conv_workspace_bytes", "sd_conv_weight_layout_tc", "sd_conv_pack_weights_simt",
    "sd_conv_pack_weights_tc", "sd_conv_lif_simt", "sd_conv_lif_tc", "sd_conv_tc_supported", "sd_debug_tc_trace", "sd_debug_tc_reload_knobs", "sd_conv_wgrad_workspace_bytes", "sd_conv_wgrad",
    "sd_bn_train_forward", "sd_bn_backward", "sd_bn_local_stats", "sd_bn_backward_reduce", "sd_bn_backward_apply", "sd_philox_uniform",
    "sd_philox_exponential", "sd_philox_offset_increment", "sd_sample_step", "sd_sample_step_dev", "sd_denoiser_input", "sd_to_uint8",
    "sd_metric_workspace_bytes", "sd_metric_mse", "sd_metric_ssim", "sd_metric_feature_stats", "sd_metric_frechet",
    "sd_metric_poly_mmd2", "sd_metric_inception_score",
]


class SdError(RuntimeError):
    """A non-zero return code from libsd_b200 (SD_ERR_CUDA / SD_ERR_NO_DEVICE / SD_ERR_UNSUPPORTED)."""


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "sd_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


TRACE_LIB_PATH = os.path.join(ROOT, "build", "libsd_b200_trace.so")


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into spiking-diffusion_b200/libsd_b200.so (in-tree, travels with gpurun).

    ``trace=True`` builds the diagnostics variant instead (-DSD_TRACE: cycle stamps and the SD_TC_DBG experiment
    paths inside conv3x3_tc_kernel) as build/libsd_b200_trace.so; select it with the environment variable
    SD_B200_LIB before the first call (tools/trace_tc.py does).  The shipped library carries none of that code."""
    if trace:
        return _build_to(TRACE_LIB_PATH, os.path.join(ROOT, "build", "trace_obj"), ["-DSD_TRACE"], verbose)
    if not force and not _stale():
        return LIB_PATH
    return _build_to(LIB_PATH, CSRC, [], verbose)


def _build_to(lib_path: str, obj_dir: str, extra_flags, verbose: bool) -> str:
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(os.path.dirname(lib_path), exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        # SD_NVCC_EXTRA: experiment macros for the diagnostics build (e.g. "-DSD_TC_SINGLE_BUF"), never for the shipped one
        more = os.environ.get("SD_NVCC_EXTRA", "").split() if extra_flags else []
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + more + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    cmd = [_nvcc(), "--shared", "-cudart", "shared", "-o", lib_path] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return lib_path


class ConvDesc(ctypes.Structure):
    """struct sd_conv_desc (include/sd_b200.h)."""
    _fields_ = [(n, ctypes.c_int) for n in
                ("T", "B", "C_in", "H_in", "W_in", "C_out", "H_out", "W_out", "kh", "kw", "stride", "pad",
                 "transposed", "in_kind", "out_kind", "in_T", "C_in0")] + \
               [("tau", ctypes.c_float), ("v_threshold", ctypes.c_float), ("v_reset", ctypes.c_float),
                ("hard_reset", ctypes.c_int), ("nsplit", ctypes.c_int), ("concurrent", ctypes.c_int)]


class ConvArgs(ctypes.Structure):
    """struct sd_conv_args (include/sd_b200.h)."""
    _fields_ = [("in_", ctypes.c_void_p), ("in2", ctypes.c_void_p), ("weights", ctypes.c_void_p),
                ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("out_sum", ctypes.c_void_p), ("memout_coef_host", ctypes.c_void_p),
                ("workspace", ctypes.c_void_p), ("in_scalar", ctypes.c_float)]


IN_REAL_CONST, IN_REAL_SEQ, IN_STF, IN_STF8, IN_TOKENS = 0, 1, 2, 3, 4
OUT_LIF, OUT_REAL_SEQ, OUT_MEMOUT_TANH, OUT_MEAN_T, OUT_LIF8, OUT_CURRENT_SEQ = 0, 1, 2, 3, 4, 5
SD_ERR_INVALID, SD_ERR_CUDA, SD_ERR_NO_DEVICE, SD_ERR_UNSUPPORTED = 1, 2, 3, 4

_lib: Optional[ctypes.CDLL] = None


def _declare(lib: ctypes.CDLL) -> None:
    i, i64, u64, f, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float, ctypes.c_void_p
    pd, pa = ctypes.POINTER(ConvDesc), ctypes.POINTER(ConvArgs)
    sig = {
        "sd_last_error": (ctypes.c_char_p, []),
        "sd_version": (i, []),
        "sd_device_info": (i, [ctypes.POINTER(i)] * 4),
        "sd_stf_guard": (i64, [i]),
        "sd_stf_rows": (i64, [i, i, i]),
        "sd_stf_bytes": (i64, [i, i, i, i, i]),
        "sd_stf_from_nchw": (i, [vp, vp, i, i, i, i, i, vp]),
        "sd_stf_to_nchw": (i, [vp, vp, i, i, i, i, i, vp]),
        "sd_stf8_from_nchw": (i, [vp, vp, i, i, i, i, i, vp]),
        "sd_stf8_to_nchw": (i, [vp, vp, i, i, i, i, i, vp]),
        "sd_stf_upsample2x": (i, [vp, vp, i, i, i, i, i, vp]),
        "sd_stf_subsample2x": (i, [vp, vp, i, i, i, i, i, vp]),
        "sd_channel_affine": (i, [vp, vp, vp, vp, i64, i, i64, vp]),
        "sd_state_convert": (i, [vp, vp, i, i, i, i, i, vp]),
        "sd_lif_forward": (i, [vp, vp, vp, vp, i, i64, f, f, f, i, i, vp]),
        "sd_lif_backward": (i, [vp, vp, vp, vp, vp, i, i64, f, f, f, i, i, i, f, vp]),
        "sd_memout": (i, [vp, vp, vp, i, i64, i, vp]),
        "sd_vq_feature": (i, [vp, vp, vp, vp, i, i, i, i, i, vp]),
        "sd_vq_lookup": (i, [vp, vp, vp, vp, i64, i, i, vp]),
        "sd_vq_gather": (i, [vp, vp, vp, i, i, i, i, i, vp]),
        "sd_conv_weight_bytes_simt": (i64, [pd]),
        "sd_conv_weight_bytes_tc": (i64, [pd]),
        "sd_conv_workspace_bytes": (i64, [pd]),
        "sd_conv_weight_layout_tc": (i64, [pd]),
        "sd_conv_pack_weights_simt": (i, [pd, vp, vp, vp]),
        "sd_conv_pack_weights_tc": (i, [pd, vp, vp, vp, vp]),
        "sd_conv_lif_simt": (i, [pd, pa, vp]),
        "sd_conv_lif_tc": (i, [pd, pa, vp]),
        "sd_conv_tc_supported": (i, [pd]),
        "sd_debug_tc_trace": (i, [vp]),
        "sd_debug_tc_reload_knobs": (i, []),
        "sd_conv_wgrad_workspace_bytes": (i64, [pd]),
        "sd_conv_wgrad": (i, [pd, vp, vp, vp, vp, vp, vp]),
        "sd_bn_train_forward": (i, [vp, vp, vp, vp, vp, vp, i64, i, i64, f, vp]),
        "sd_bn_backward": (i, [vp, vp, vp, vp, vp, vp, vp, vp, i64, i, i64, f, vp]),
        "sd_bn_local_stats": (i, [vp, vp, vp, i64, i, i64, vp]),
        "sd_bn_backward_reduce": (i, [vp, vp, vp, vp, vp, vp, i64, i, i64, f, vp]),
        "sd_bn_backward_apply": (i, [vp, vp, vp, vp, vp, vp, vp, vp, i64, i, i64, f, vp]),
        "sd_philox_uniform": (i, [vp, i64, u64, u64, i64, i64, ctypes.POINTER(u64), vp]),
        "sd_philox_exponential": (i, [vp, i64, u64, u64, i64, i64, ctypes.POINTER(u64), vp]),
        "sd_philox_offset_increment": (i, [i64, ctypes.POINTER(u64)]),
        "sd_sample_step": (i, [vp, vp, vp, vp, i64, i, i, f, u64, u64, u64, i64, i64, vp]),
        "sd_sample_step_dev": (i, [vp, vp, vp, vp, i64, i, i, f, vp, u64, u64, i64, i64, vp]),
        "sd_denoiser_input": (i, [vp, vp, i, i, i, i, vp]),
        "sd_to_uint8": (i, [vp, vp, i64, vp]),
        "sd_metric_workspace_bytes": (i64, [i64, i, i]),
        "sd_metric_mse": (i, [vp, vp, i64, vp, vp, vp]),
        "sd_metric_ssim": (i, [vp, vp, i, i, i, i, i, vp, vp, vp, vp, vp]),
        "sd_metric_feature_stats": (i, [vp, i, i64, i, vp, vp, vp]),
        "sd_metric_frechet": (i, [vp, vp, vp, vp, i, vp, ctypes.POINTER(i), vp, vp]),
        "sd_metric_poly_mmd2": (i, [vp, vp, i, i, i, ctypes.c_double, ctypes.c_double, vp, vp, vp]),
        "sd_metric_inception_score": (i, [vp, i64, i, i, vp, vp, vp, vp]),
    }
    assert set(sig) == set(SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


def lib() -> ctypes.CDLL:
    """The loaded C-ABI library.  Raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        path = os.environ.get("SD_B200_LIB") or LIB_PATH   # SD_B200_LIB: the -DSD_TRACE diagnostics build (tools/)
        if not os.path.exists(path):
            raise SdError(f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU or PyTorch fallback for the CUDA path)")
        _lib = ctypes.CDLL(path)
        _declare(_lib)
    return _lib


def check(rc: int) -> None:
    """Map a C-ABI return code onto the exception type the reference raises on the same condition:
    SD_ERR_INVALID -> ValueError (SJ/activation_based/layer.py:169-170, neuron.py:92-94,707), others -> SdError."""
    if rc == 0:
        return
    msg = lib().sd_last_error().decode("utf-8", "replace")
    if rc == SD_ERR_INVALID:
        raise ValueError(msg)
    raise SdError(f"libsd_b200 error {rc}: {msg}")


# ---- device discipline ---------------------------------------------------------------------------------------------
# The C-ABI takes raw pointers and a raw stream and launches on the CUDA runtime's CURRENT device.  Every call site
# builds its arguments with ptr(tensor) ... stream_ptr() (the stream is always the last argument), so the two helpers
# also enforce that all tensors of one call live on one device and that this device is the current one; the public
# entry points (module forward methods, plans) run under ``on_device_of`` so that a model moved with ``.to('cuda:1')``
# works without the caller touching torch.cuda.set_device.
import threading

_tls = threading.local()


def ptr(t) -> Optional[int]:
    if t is None:
        return None
    d = t.get_device()          # -1 for a CPU tensor (host-side coefficient arrays never come through here)
    if d >= 0:
        seen = getattr(_tls, "devices", None)
        if seen is None:
            seen = _tls.devices = set()
        seen.add(d)
    return t.data_ptr()


def stream_ptr() -> int:
    import torch
    seen = getattr(_tls, "devices", None)
    cur = torch._C._cuda_getDevice()
    if seen:
        if len(seen) > 1 or cur not in seen:
            devs = sorted(seen)
            seen.clear()
            if len(devs) > 1:
                raise ValueError(f"all tensors of one libsd_b200 call must live on one CUDA device, got cuda:{devs}")
            raise RuntimeError(f"tensors live on cuda:{devs[0]} but the current CUDA device is cuda:{cur}; "
                               "run the call under torch.cuda.device(tensor.device) (the public entry points do)")
        seen.clear()
    return torch._C._cuda_getCurrentRawStream(cur)


def first_cuda_device(*objs):
    """Device of the first CUDA tensor found among tensors / modules / sequences of them, or None."""
    import torch
    for o in objs:
        if isinstance(o, torch.Tensor):
            if o.is_cuda:
                return o.device
        elif isinstance(o, torch.nn.Module):
            for t in o.parameters():
                if t.is_cuda:
                    return t.device
                break
            for t in o.buffers():
                if t.is_cuda:
                    return t.device
                break
        elif isinstance(o, (list, tuple)):
            d = first_cuda_device(*o)
            if d is not None:
                return d
    return None


def on_device_of(fn):
    """Decorator for public entry points: run ``fn`` with the CUDA device of its first CUDA tensor / module argument
    current (torch.cuda.device), so kernels, streams and allocations all land on the data's device."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        import torch
        seen = getattr(_tls, "devices", None)
        if seen:
            seen.clear()        # nothing may be left over from a call that raised between ptr() and stream_ptr()
        dev = first_cuda_device(*args, *kwargs.values())
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper
