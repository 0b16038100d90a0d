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
    """reference samplers.py:104-177 -- buffered rejection sampler with an empirical index code.  Host-side logic (torch /
    NumPy float64), candidates regenerated on the GPU; see rejection_sampling.py for the two places where the
    reference's unseeded draws are replaced by seeded ones."""

    def __init__(self, sample_buffer_size, r_buffer_size, use_pseudo_sampler=False, name="rejection_sampler", **kwargs):
        super().__init__(name=name, **kwargs)
        from rec.coding.sample_generator import NaiveSampleGenerator, PseudoSampleGenerator
        self.sample_buffer_size = int(sample_buffer_size)
        self.r_buffer_size = int(r_buffer_size)
        self.sample_generator = (PseudoSampleGenerator if use_pseudo_sampler else NaiveSampleGenerator)(self.sample_buffer_size)
        self.average_count = 0.
        self._initialized = False
        self.acceptance_probabilities = np.zeros(self.r_buffer_size, dtype=np.float64)
        self.spillover_probability = 0.
        self.spillover_acceptance_probability = 0.

    def update(self, target, coder, seed=0):
        """running average of the per-index acceptance probabilities (reference :136-149)"""
        from rec.coding.rejection_sampling import get_r_pstar, get_t_p_mass
        log_ratios, t_mass, p_mass = get_t_p_mass(target, coder, seed=seed + int(self.average_count))
        _, pstar = get_r_pstar(log_ratios, t_mass, p_mass, self.r_buffer_size, dtype=np.float64)
        acc = pstar - np.concatenate(([0.], pstar[:-1]))
        self.acceptance_probabilities = (self.acceptance_probabilities * self.average_count + acc) / (self.average_count + 1.)
        self.average_count += 1.
        self.spillover_probability = 1. - float(np.sum(self.acceptance_probabilities))
        self.spillover_acceptance_probability = float(self.acceptance_probabilities[-1] /
                                                      (1. - np.sum(self.acceptance_probabilities[:-1])))
        self._initialized = True

    def get_codelength(self, index):
        """nats (reference :151-159)"""
        assert self._initialized
        index = int(index)
        if index < self.r_buffer_size:
            return float(-np.log(self.acceptance_probabilities[index]))
        return float(-(np.log(self.spillover_probability) +
                       np.log(1. - self.spillover_acceptance_probability) * (index - self.r_buffer_size) +
                       np.log(self.spillover_acceptance_probability)))

    def coded_sample(self, target, coder, seed):
        from rec.coding.rejection_sampling import gaussian_rejection_sample_small
        return gaussian_rejection_sample_small(t_dist=target, p_dist=coder, sample_buffer_size=self.sample_buffer_size,
                                               r_buffer_size=self.r_buffer_size, sample_generator=self.sample_generator,
                                               seed=int(seed))

    def decode_sample(self, coder, sample_index, seed):
        sample_index = int(sample_index)
        return self.sample_generator.generate_index(sample_index % self.sample_buffer_size, coder,
                                                    seed=int(seed) + sample_index // self.sample_buffer_size)
