"""rec.coding.sample_generator -- candidate buffers of the rejection sampler
(reference: rec/coding/sample_generator.py:7-66, NaiveSampleGenerator).

`coder.sample((buffer,), seed=seed)` after `tf.random.set_seed(seed)` is the float32 tf.random.normal stream with
(global seed, op seed) = (seed, seed), times scale plus loc; the stream is regenerated on the GPU by
`irec_normal_stream_seeded` (Philox4x32-10 + Box-Muller, bit-compatible restatement), so encoder and decoder see the
same candidates.  The PseudoSampleGenerator (reference :69-133) is not built."""
import abc
import math

import torch

from irec_b200 import engine as E

_LOG_SQRT_2PI = 0.5 * math.log(2. * math.pi)


def normal_log_prob(x, loc, scale):
    """TFP 0.9 Normal._log_prob: -0.5 * squared_difference(x / s, m / s) - (0.5 log 2pi + log s)"""
    return -0.5 * (x / scale - loc / scale) ** 2 - (_LOG_SQRT_2PI + torch.log(scale))


class SampleGenerator(abc.ABC):
    @abc.abstractmethod
    def get_ratios(self, target, coder, seed):
        """fills the buffer with samples of the coder distribution -> log likelihood ratios [buffer]"""

    @abc.abstractmethod
    def get_index(self, i):
        """sample i of the buffer"""

    @abc.abstractmethod
    def generate_index(self, i, coder, seed):
        """regenerates the buffer of `seed` and returns its sample i (decoding)"""


class NaiveSampleGenerator(SampleGenerator):
    def __init__(self, sample_buffer_size, **kwargs):
        self.sample_buffer_size = int(sample_buffer_size)
        self.samples = None

    def _buffer(self, coder, seed, first=0, count=None):
        loc, scale = E._f32c(coder.loc, "cuda"), E._f32c(coder.scale, "cuda")
        count = self.sample_buffer_size if count is None else count
        d = loc.numel()
        z = E.normal_stream_seeded(seed, seed, first * d, count * d, device=loc.device)
        return z.reshape((count,) + tuple(loc.shape)) * scale + loc

    def get_ratios(self, target, coder, seed):
        self.samples = self._buffer(coder, seed)
        t_loc, t_scale = E._f32c(target.loc, "cuda"), E._f32c(target.scale, "cuda")
        c_loc, c_scale = E._f32c(coder.loc, "cuda"), E._f32c(coder.scale, "cuda")
        diff = normal_log_prob(self.samples, t_loc, t_scale) - normal_log_prob(self.samples, c_loc, c_scale)
        return diff.reshape(self.sample_buffer_size, -1).sum(dim=1)

    def get_index(self, i):
        return self.samples[int(i)]

    def generate_index(self, i, coder, seed):
        return self._buffer(coder, seed, first=int(i), count=1)[0]      # counter-based: only the asked sample is generated
