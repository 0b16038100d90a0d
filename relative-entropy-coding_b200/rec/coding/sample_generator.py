"""rec.coding.sample_generator -- candidate buffers of the rejection sampler
(reference: rec/coding/sample_generator.py:7-66, NaiveSampleGenerator).

`coder.sample((buffer,), seed=seed)` after `tf.random.set_seed(seed)` is the float32 tf.random.normal stream with
(global seed, op seed) = (seed, seed), times scale plus loc; the stream is regenerated on the GPU by
`irec_normal_stream_seeded` (Philox4x32-10 + Box-Muller, bit-compatible restatement), so encoder and decoder see the
same candidates.  PseudoSampleGenerator (reference :69-133): a few true samples recombined per random group of dims."""
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


class PseudoSampleGenerator(SampleGenerator):
    """reference :69-133 -- `n_true_samples` real draws of the coder; the dims are thrown into `n_groups` random groups and
    pseudo-sample i takes, for every group, the values of one randomly assigned true sample.  All three random tensors come
    from the SAME Philox stream (every op is called with seed=seed after set_seed(seed)), restated by
    irec_normal_stream_seeded / irec_uniform_int_stream."""

    def __init__(self, sample_buffer_size, n_true_samples=50, n_groups=50, **kwargs):
        self.sample_buffer_size = int(sample_buffer_size)
        self.n_true_samples = int(n_true_samples)
        self.n_groups = int(n_groups)
        self.samples = None
        self.group_assignments = None
        self.sample_assignments = None

    def _draw(self, coder, seed):
        loc, scale = E._f32c(coder.loc, "cuda"), E._f32c(coder.scale, "cuda")
        d = loc.numel()
        z = E.normal_stream_seeded(seed, seed, 0, self.n_true_samples * d, device=loc.device)
        self.samples = z.reshape((self.n_true_samples,) + tuple(loc.shape)) * scale + loc
        self.group_assignments = E.uniform_int_stream(seed, seed, 0, self.n_groups, 0, d, device=loc.device).long()      # [d]
        self.sample_assignments = E.uniform_int_stream(seed, seed, 0, self.n_true_samples, 0,
                                                       self.n_groups * self.sample_buffer_size,
                                                       device=loc.device).long().reshape(self.n_groups, self.sample_buffer_size)
        return loc, scale

    def get_ratios(self, target, coder, seed):
        loc, scale = self._draw(coder, seed)
        t_loc, t_scale = E._f32c(target.loc, "cuda"), E._f32c(target.scale, "cuda")
        ratios = (normal_log_prob(self.samples, t_loc, t_scale) - normal_log_prob(self.samples, loc, scale))
        flat = ratios.reshape(self.n_true_samples, -1)                                          # [n_true, d]
        mask = torch.nn.functional.one_hot(self.group_assignments, self.n_groups).to(flat.dtype)  # [d, n_groups]
        group_ratios = flat @ mask                                                              # [n_true, n_groups]
        # ratio of pseudo-sample i = sum_g group_ratios[sample_assignments[g, i], g]
        g = torch.arange(self.n_groups, device=flat.device)[:, None]
        return group_ratios[self.sample_assignments, g].sum(dim=0)

    def _assemble(self, i):
        group_indices = self.sample_assignments[:, int(i)]                    # [n_groups] true-sample index per group
        per_dim = group_indices[self.group_assignments]                       # [d]
        flat = self.samples.reshape(self.n_true_samples, -1)
        d = torch.arange(flat.shape[1], device=flat.device)
        return flat[per_dim, d].reshape(self.samples.shape[1:])

    def get_index(self, i):
        return self._assemble(i)

    def generate_index(self, i, coder, seed):
        self._draw(coder, seed)
        return self._assemble(i)
