"""Import the UNMODIFIED reference (R/snn_model + the vendored SpikingJelly zip) -- TEST / BENCH INFRASTRUCTURE.

Two sources, in this order: the install under the git-ignored ``baseline/_ref`` made by ``oracle/install_ref.py``
(unmodified copies; it travels to the GPU box), else /root/reference itself (this build container only).  It is used
to (a) prove ``oracle/snn_oracle.py`` bit-identical to the reference at the reference's hard-coded T=16 / 7x7,
(b) generate the committed golden vectors (``oracle/gen_golden.py``), (c) give ``bench.py --impl reference`` the
reference's own code to time on the host CPU, and (d) provide the reference's metric functions to the metric parity
tests.  Nothing under ``spiking-diffusion_b200/`` imports it.

Procedure: the zip has no top-level package directory, so it is extracted into a
directory *named* ``spikingjelly``; ``matplotlib`` and ``spikingjelly.visualizing`` are imported at module
top by the reference (R/snn_model/vae_model.py:17-18) but absent here, so empty stubs are injected.
No reference source is copied into the repository: the extraction goes to a temp directory.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
import zipfile

REF_ROOT = "/root/reference/Spiking-Diffusion-release"
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED_ROOT = os.path.join(_ROOT, "baseline", "_ref")


def _installed() -> bool:
    return (os.path.isfile(os.path.join(INSTALLED_ROOT, "INSTALLED.json"))
            and os.path.isfile(os.path.join(INSTALLED_ROOT, "spikingjelly", "__init__.py")))


def available() -> bool:
    return _installed() or os.path.isfile(os.path.join(REF_ROOT, "spikingjelly.zip"))


def source() -> str:
    """Where load() takes the reference from: 'baseline/_ref' (install) or '/root/reference' (mount)."""
    return "baseline/_ref" if _installed() else "/root/reference"


_loaded = None


def load():
    """Returns a namespace with the reference's classes: SNN_VQVAE, VectorQuantizer, Encoder, Decoder,
    DummyModel, AbsorbingDiffusion, MembraneOutputLayer, PSP, and SJ's neuron / functional / layer / surrogate."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference neither installed under {INSTALLED_ROOT} (oracle/install_ref.py) nor mounted at {REF_ROOT}")
    sys.dont_write_bytecode = True
    if _installed():
        dst, rel_root = INSTALLED_ROOT, os.path.join(INSTALLED_ROOT, "Spiking-Diffusion-release")
    else:
        dst, rel_root = os.path.join(tempfile.gettempdir(), "sd_ref_sj"), REF_ROOT
        pkg = os.path.join(dst, "spikingjelly")
        if not os.path.isfile(os.path.join(pkg, "__init__.py")):
            os.makedirs(pkg, exist_ok=True)
            zipfile.ZipFile(os.path.join(REF_ROOT, "spikingjelly.zip")).extractall(pkg)
    for name in ("matplotlib", "matplotlib.pyplot", "spikingjelly.visualizing"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path[:0] = [dst, rel_root]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from spikingjelly.activation_based import neuron, functional, layer, surrogate
        import spikingjelly
        spikingjelly.visualizing = sys.modules["spikingjelly.visualizing"]
        from snn_model import vae_model, vq_diffusion, snn_layers
    ns = types.SimpleNamespace(
        neuron=neuron, functional=functional, layer=layer, surrogate=surrogate,
        SNN_VQVAE=vae_model.SNN_VQVAE, VectorQuantizer=vae_model.VectorQuantizer,
        Encoder=vae_model.Encoder, Decoder=vae_model.Decoder,
        DummyModel=vq_diffusion.DummyModel, AbsorbingDiffusion=vq_diffusion.AbsorbingDiffusion,
        MembraneOutputLayer=snn_layers.MembraneOutputLayer, PSP=snn_layers.PSP,
        vq_diffusion=vq_diffusion, vae_model=vae_model, root=rel_root)
    _loaded = ns
    return ns


def load_metrics():
    """The reference's metric code that needs no pretrained weights: ``pytorch_ssim`` (R/metric/pytorch_ssim/__init__.py)
    and ``calculate_frechet_distance`` / ``sqrtm`` of R/metric/Fid_score.py:14-17,116-173.  Fid_score.py imports
    torchvision's inception_v3 at module top; only the two pure-numpy functions are used."""
    ns = load()
    import importlib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ssim_mod = importlib.import_module("metric.pytorch_ssim")
        fid_mod = importlib.import_module("metric.Fid_score")
    return types.SimpleNamespace(pytorch_ssim=ssim_mod, calculate_frechet_distance=fid_mod.calculate_frechet_distance,
                                 sqrtm=fid_mod.sqrtm, ref=ns)


class redirect_cuda_to_cpu:
    """Context manager: the reference's sample() hard-codes ``device = 'cuda'`` (R/snn_model/vq_diffusion.py:105,29).
    To time its CPU implementation unmodified, the name ``torch`` in that module's namespace is replaced by a proxy
    whose tensor factories rewrite ``device='cuda'`` to ``'cpu'``; every other attribute is torch's own."""

    FACTORIES = ("ones", "zeros", "full", "zeros_like", "ones_like", "rand_like", "empty", "tensor", "arange")

    def __init__(self, module):
        self.module = module

    def __enter__(self):
        import torch

        class _Proxy:
            def __getattr__(self, name):
                attr = getattr(torch, name)
                if name in redirect_cuda_to_cpu.FACTORIES:
                    def wrapped(*a, **k):
                        if str(k.get("device", "")).startswith("cuda"):
                            k["device"] = "cpu"
                        return attr(*a, **k)
                    return wrapped
                return attr
        self._saved = self.module.torch
        self.module.torch = _Proxy()
        return self

    def __exit__(self, *exc):
        self.module.torch = self._saved
        return False
