"""spiking_diffusion_b200 -- B200-native (sm_100a) implementation of the Spiking-Diffusion hot path.

Drop-in surface (same names and signatures as the reference):

    from spiking_diffusion_b200.activation_based import neuron, functional, layer, surrogate
    from spiking_diffusion_b200.snn_model.vae_model import SNN_VQVAE, VectorQuantizer, Encoder, Decoder
    from spiking_diffusion_b200.snn_model.vq_diffusion import DummyModel, AbsorbingDiffusion

Every forward call lands in one C-ABI symbol of ``libsd_b200.so`` (include/sd_b200.h); there is no CPU or
PyTorch fallback.  The directory is named ``spiking-diffusion_b200``; ``spiking_diffusion_b200.py`` at the
repository root makes it importable under a valid Python identifier.
"""
from . import _lib  # noqa: F401
from ._lib import SdError, build  # noqa: F401

__all__ = ["SdError", "build", "activation_based", "snn_model", "engine", "synth"]
