"""Import the UNMODIFIED reference (R/snn_model + the vendored SpikingJelly zip) -- TEST INFRASTRUCTURE.

Only usable where /root/reference is mounted (this build container).  It is used to (a) prove
``oracle/snn_oracle.py`` bit-identical to the reference at the reference's hard-coded T=16 / 7x7, and
(b) generate the committed golden vectors (``oracle/gen_golden.py``).  Nothing on the GPU box imports it.

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


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "spikingjelly.zip"))


_loaded = None


def load():
    """Returns a namespace with the reference's classes: SNN_VQVAE, VectorQuantizer, Encoder, Decoder,
    DummyModel, AbsorbingDiffusion, MembraneOutputLayer, PSP, and SJ's neuron / functional / layer / surrogate."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference not mounted at " + REF_ROOT)
    sys.dont_write_bytecode = True
    dst = os.path.join(tempfile.gettempdir(), "sd_ref_sj")
    pkg = os.path.join(dst, "spikingjelly")
    if not os.path.isfile(os.path.join(pkg, "__init__.py")):
        os.makedirs(pkg, exist_ok=True)
        zipfile.ZipFile(os.path.join(REF_ROOT, "spikingjelly.zip")).extractall(pkg)
    for name in ("matplotlib", "matplotlib.pyplot", "spikingjelly.visualizing"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path[:0] = [dst, REF_ROOT]
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
        MembraneOutputLayer=snn_layers.MembraneOutputLayer, PSP=snn_layers.PSP)
    _loaded = ns
    return ns
