"""rec.coding.samplers -- sampler plug-in contract and the importance sampler
(reference: rec/coding/samplers.py:16-101)."""
import abc

import numpy as np
import torch

from irec_b200 import engine as E
from rec.coding.utils import CodingError
from rec.coding.importance_sampling import (encode_gaussian_importance_sample, decode_gaussian_importance_sample,
                                            importance_num_samples)


class Sampler(abc.ABC):
    """reference samplers.py:16-58.  Unknown keyword arguments are accepted and ignored: the reference's
    examples pass `extrapolate_auxiliary_vars` through **sampler_args (compression_performance.py:60-77)."""

    def __init__(self, name="sampler", **kwargs):
        self.name = name

    @abc.abstractmethod
    def coded_sample(self, target, coder, seed):
        """-> (sample index, sample)"""

    @abc.abstractmethod
    def decode_sample(self, coder, sample_index, seed):
        """-> sample with the given index"""

    @abc.abstractmethod
    def get_codelength(self, index):
        pass

    @abc.abstractmethod
    def update(self, target, coder):
        pass


class ImportanceSampler(Sampler):
    """reference samplers.py:61-101"""

    def __init__(self, coding_bits, alpha=np.inf, name="importance_sampler", **kwargs):
        super().__init__(name=name, **kwargs)
        self.alpha = alpha
        self.coding_bits = coding_bits

    @property
    def n_samples(self):
        return importance_num_samples(self.coding_bits)

    def coded_sample(self, target, coder, seed):
        return encode_gaussian_importance_sample(t_loc=target.loc, t_scale=target.scale, p_loc=coder.loc,
                                                 p_scale=coder.scale, coding_bits=self.coding_bits, seed=seed,
                                                 alpha=self.alpha)

    def decode_sample(self, coder, sample_index, seed):
        return decode_gaussian_importance_sample(p_loc=coder.loc, p_scale=coder.scale, index=sample_index, seed=seed)

    def update(self, target, coder):
        pass    # "ImportanceSampler doesn't require updating!" (reference samplers.py:98)

    def get_codelength(self, index):
        return float(np.float32(self.coding_bits) * np.float32(np.log(2.)))


class RejectionSampler(Sampler):
    """Out of scope of the B200 hot path (reference samplers.py:104-177,
    rejection_sampling.py, sample_generator.py).  Present so that imports of the reference's names work."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("RejectionSampler is outside the accelerated iREC path; use ImportanceSampler or "
                                  "BeamSearchCoder")

    def coded_sample(self, target, coder, seed):
        raise NotImplementedError

    def decode_sample(self, coder, sample_index, seed):
        raise NotImplementedError

    def get_codelength(self, index):
        raise NotImplementedError

    def update(self, target, coder):
        raise NotImplementedError
