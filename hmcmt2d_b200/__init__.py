"""hmcmt2d_b200 — B200-native (sm_100a) drop-in for the forward + adjoint-gradient hot path of
CUG-EMI/HMCMT2D.  Host-side mirror of the reference's solver / sampler interface above the C ABI
of libhmcmt_b200.so.  No CPU fallback: importing the compute entry points without the built CUDA
library raises."""
from . import lib  # noqa: F401

__all__ = ["lib"]
