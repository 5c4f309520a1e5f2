"""Mirror of the reference's ``snn_model`` package (R/snn_model/) for the hot-path classes."""
from .snn_layers import PSP, MembraneOutputLayer  # noqa: F401
from .vae_model import Decoder, Encoder, SNN_VQVAE, VectorQuantizer  # noqa: F401
from .vq_diffusion import AbsorbingDiffusion, DummyModel  # noqa: F401
