"""Host-side mirror of the reference's src/nr/network interface for the hot path (same class names, ctor cfg handling,
forward(data) signature, output keys and state_dict keys), with the hot path routed to the CUDA kernels."""
from .renderer import NeuralRayRenderer, GraspNeRF, name2network  # noqa: F401

from .config import NRVGN_SDF_CFG  # noqa: E402,F401
