"""rec.coding.importance_sampling -- one-partition importance sampler
(reference: rec/coding/importance_sampling.py:9-103), alpha = inf branch on the GPU."""
import math

import numpy as np
import torch

from irec_b200 import engine as E
from rec.coding.utils import CodingError

_DEVICE = "cuda"


def importance_num_samples(coding_bits):
    """S = int32(ceil(exp(coding_bits * log(2.))))  (reference :51), with the float contract of DESIGN.md:
    float32 product, exp evaluated in float64 and rounded once to float32, then ceil."""
    prod = np.float32(coding_bits) * np.float32(0.6931471805599453)
    e = np.float32(math.exp(float(prod)))
    return int(math.ceil(float(e)))


def encode_gaussian_importance_sample(t_loc, t_scale, p_loc, p_scale, coding_bits, seed, log_weighting_fn=None,
                                      alpha=float('inf')):
    """-> (index, sample); index is a 0-d int64 CPU tensor (host-visible like the reference's eager tensor)."""
    if alpha < 1.:
        raise CodingError(f"Alpha must be in the range [1, inf), but {alpha} was given!")
    if not math.isinf(alpha):
        raise CodingError("finite alpha is not supported: the reference's Gumbel helper (rec/coding/utils.py:10-12) "
                          "takes the log of a normal draw and produces NaNs; every shipped config uses alpha=inf")
    if log_weighting_fn is not None:
        raise CodingError("log_weighting_fn is not supported by the fused kernel")
    tl, ts = E._f32c(t_loc, _DEVICE), E._f32c(t_scale, _DEVICE)
    pl, ps = E._f32c(p_loc, _DEVICE), E._f32c(p_scale, _DEVICE)
    shape = tl.shape
    index, sample = E.is_coded_sample(tl.reshape(-1), ts.reshape(-1), pl.reshape(-1), ps.reshape(-1),
                                      importance_num_samples(coding_bits), seed)
    return index.cpu().reshape(()), sample.reshape(shape)


def decode_gaussian_importance_sample(p_loc, p_scale, index, seed):
    pl, ps = E._f32c(p_loc, _DEVICE), E._f32c(p_scale, _DEVICE)
    shape = pl.shape
    return E.is_decode_sample(pl.reshape(-1), ps.reshape(-1), int(index), seed).reshape(shape)
