"""Mirror of ``spikingjelly.activation_based`` restricted to what Spiking-Diffusion instantiates."""
from . import base, surrogate, neuron, functional, layer  # noqa: F401
