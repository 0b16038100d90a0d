"""rec.coding.utils -- error type of the coding package (reference: rec/coding/utils.py:4-7).

`stateless_gumbel_sample` (reference utils.py:10-12) is only reached by the alpha < inf branch of the
importance sampler, which feeds a *normal* draw into -log(-log(.)) and therefore yields NaNs; no
shipped configuration uses it (alpha = inf everywhere).  It is deliberately not provided; the
importance sampler raises CodingError for finite alpha.
"""
from irec_b200.engine import CodingError

__all__ = ["CodingError"]
