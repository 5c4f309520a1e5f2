"""The C-ABI library loads on a CPU-only box, exports exactly what include/sd_b200.h declares, and refuses to
compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

import spiking_diffusion_b200 as sd
from spiking_diffusion_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in header_symbols():
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (sd_[a-z0-9_]+)$", out, flags=re.M))
    assert exported == set(header_symbols())


def test_struct_layout_matches_header(tmp_path):
    """The ctypes mirrors of sd_conv_desc / sd_conv_args against what a C compiler makes of include/sd_b200.h."""
    import subprocess
    # 17 ints + 3 floats + 3 ints, no padding; 10 pointers + one float (padded to pointer alignment)
    assert ctypes.sizeof(_lib.ConvDesc) == 23 * 4
    assert ctypes.sizeof(_lib.ConvArgs) == 11 * ctypes.sizeof(ctypes.c_void_p)
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sd_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(sd_conv_desc), sizeof(sd_conv_args), '
                   'offsetof(sd_conv_desc, tau), offsetof(sd_conv_desc, concurrent), offsetof(sd_conv_args, workspace), '
                   'offsetof(sd_conv_args, in_scalar)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_lib.ConvDesc), ctypes.sizeof(_lib.ConvArgs), _lib.ConvDesc.tau.offset, _lib.ConvDesc.concurrent.offset,
            _lib.ConvArgs.workspace.offset, _lib.ConvArgs.in_scalar.offset]
    assert got == want, (got, want)


def test_geometry_queries_are_host_only():
    L = _lib.lib()
    assert L.sd_version() == 1
    assert L.sd_stf_guard(7) == 8 and L.sd_stf_guard(28) == 32
    # 256 images x 49 pixels = 12544 rows = 98 tiles of 128 (dense: no padding is stored), + 2 guards of 8
    assert L.sd_stf_rows(256, 7, 7) == 12544 + 16
    assert L.sd_stf_bytes(4, 256, 64, 7, 7) == 4 * 8 * (12544 + 16) * 8 * 2
    assert L.sd_stf_bytes(1, 1, 3, 7, 7) == 1 * 1 * (128 + 16) * 8 * 2  # C rounded up to 8, rows to 128


def test_argument_errors_mirror_reference_exceptions():
    L = _lib.lib()
    # LIFNode asserts tau > 1 (SJ/activation_based/neuron.py:707) -> SD_ERR_INVALID -> ValueError
    rc = L.sd_lif_forward(None, None, None, None, 4, 16, 1.0, 1.0, 0.0, 1, 1, None)
    assert rc == _lib.SD_ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert b"tau" in L.sd_last_error()
    # empty input is a no-op, like the reference on an empty tensor
    assert L.sd_lif_forward(None, None, None, None, 0, 0, 2.0, 1.0, 0.0, 1, 1, None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    L = _lib.lib()
    x = torch.zeros(4, 16)
    rc = L.sd_lif_forward(x.data_ptr(), x[0].data_ptr(), x.data_ptr(), None, 4, 16, 2.0, 1.0, 0.0, 1, 1, None)
    assert rc == _lib.SD_ERR_NO_DEVICE
    with pytest.raises(sd.SdError):
        _lib.check(rc)
    from spiking_diffusion_b200.activation_based import neuron
    n = neuron.LIFNode(step_mode="m").eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        n(torch.zeros(4, 8))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import or execute it."""
    pkg = os.path.join(ROOT, "spiking-diffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "snn_oracle" not in text, f


def test_argument_checks_of_the_layout_and_training_entry_points():
    """Bad arguments are rejected on the host, before any device work (SD_ERR_INVALID -> ValueError), with a message."""
    import ctypes
    L = _lib.lib()
    buf = torch.zeros(64)
    p = buf.data_ptr()
    for rc in (L.sd_stf_upsample2x(p, p, 4, 1, 8, 7, 7, None),            # aliased in / out
               L.sd_stf_upsample2x(None, p, 4, 1, 8, 7, 7, None),         # null input
               L.sd_stf_subsample2x(p, p + 16, 0, 1, 8, 7, 7, None),      # T = 0
               L.sd_stf_subsample2x(p + 4, p + 32, 1, 1, 8, 2, 2, None),  # misaligned
               L.sd_bn_local_stats(p, p, None, 1, 1, 1, None),            # null output
               L.sd_bn_backward_reduce(p, p, p, p, p, p, 0, 1, 1, 1e-5, None),
               L.sd_bn_backward_apply(p, p, p, p, None, None, p, p, 1, 1, 1, 1e-5, None),
               L.sd_sample_step(p, p, p, None, 4, 2000, 1, 1.0, 0, 0, 0, 0, 4, None),   # K beyond 1024
               L.sd_sample_step(p, p, p, None, 4, 128, 0, 1.0, 0, 0, 0, 0, 4, None),    # t = 0
               L.sd_vq_lookup(p, p, p, None, 4, 0, 8, None)):                           # D = 0
        assert rc == _lib.SD_ERR_INVALID, L.sd_last_error()
        assert len(L.sd_last_error()) > 0
        with pytest.raises(ValueError):
            _lib.check(rc)
    d = _lib.ConvDesc()
    assert L.sd_conv_wgrad_workspace_bytes(ctypes.byref(d)) == 0           # invalid descriptor: nothing to allocate
    assert L.sd_conv_wgrad(ctypes.byref(d), p, p, p, None, p, None) == _lib.SD_ERR_INVALID
    d.T, d.B, d.C_in, d.H_in, d.W_in, d.C_out, d.H_out, d.W_out = 2, 3, 8, 7, 7, 16, 7, 7
    d.kh = d.kw = 3
    d.stride = d.pad = 1
    d.in_kind, d.out_kind, d.in_T, d.C_in0 = _lib.IN_REAL_SEQ, _lib.OUT_REAL_SEQ, 2, 8
    d.tau, d.v_threshold, d.hard_reset, d.nsplit = 2.0, 1.0, 1, 2
    ws = L.sd_conv_wgrad_workspace_bytes(ctypes.byref(d))
    assert ws >= 9 * 8 * 16 * 4 and ws % (9 * 8 * 16 * 4) == 0             # splits x taps x C_in x C_out floats
    assert L.sd_conv_wgrad(ctypes.byref(d), p, p, p, None, None, None) == _lib.SD_ERR_INVALID   # workspace missing
