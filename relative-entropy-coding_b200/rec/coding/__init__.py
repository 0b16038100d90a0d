"""Drop-in for the reference's `rec.coding` package (reference: rec/coding/__init__.py:1-2),
backed by hand-written sm_100a CUDA kernels (libirec.so).  No TensorFlow, no CPU fallback."""
from rec.coding.coder import Coder, GaussianCoder
from rec.coding.beam_search_coder import BeamSearchCoder

__all__ = ["Coder", "GaussianCoder", "BeamSearchCoder"]
