"""Mirror of the reference's ``snn_model`` package (R/snn_model/) for the hot-path classes."""
from .snn_layers import PSP, MembraneOutputLayer  # noqa: F401
from .vae_model import Decoder, Encoder, SNN_VQVAE, VectorQuantizer  # noqa: F401
from .vq_diffusion import AbsorbingDiffusion, DummyModel, get_data_for_diff, load_reference_state_dict  # noqa: F401
