"""rec.coding.rejection_sampling -- the reference's buffered Gaussian rejection sampler
(reference: rec/coding/rejection_sampling.py:11-116).

Host logic (torch + NumPy float64), not part of the accelerated path; the candidate buffers come from the GPU
(sample_generator.py).  Two places of the reference draw from TensorFlow's GLOBAL random state -- `t.sample(...)` in
get_t_p_mass (:13) and the acceptance uniforms (:86-87) -- so its output is not a function of (inputs, seed).  Here both
are drawn from torch Generators seeded from `seed`, which makes encode reproducible; decode never needed them."""
import math

import numpy as np
import torch

from irec_b200 import engine as E
from rec.coding.sample_generator import normal_log_prob
from rec.coding.utils import CodingError


def get_t_p_mass(t, p, n_samples=100, oversampling=100, seed=0):
    """reference :11-24 -- `n_samples` strata of the target by log-ratio, from n_samples * oversampling draws"""
    t_loc, t_scale = E._f32c(t.loc, "cuda"), E._f32c(t.scale, "cuda")
    p_loc, p_scale = E._f32c(p.loc, "cuda"), E._f32c(p.scale, "cuda")
    n = n_samples * oversampling
    gen = torch.Generator(device=t_loc.device)
    gen.manual_seed(int(seed) & 0x7fffffffffffffff)
    y = t_loc + t_scale * torch.randn((n,) + tuple(t_loc.shape), generator=gen, device=t_loc.device)
    t_mass = torch.full((n,), -math.log(n_samples), dtype=torch.float32, device=t_loc.device)
    p_mass = -math.log(n_samples) + (normal_log_prob(y, p_loc, p_scale) - normal_log_prob(y, t_loc, t_scale)).reshape(n, -1).sum(dim=1)
    log_ratios = t_mass - p_mass
    ind = torch.argsort(log_ratios)
    reduced = ind[oversampling // 2::oversampling]
    return log_ratios[reduced], t_mass[reduced], p_mass[reduced]


def _cumulative_logsumexp(x):
    return np.logaddexp.accumulate(x)


def get_r_pstar(log_ratios, t_mass, p_mass, r_buffer_size, dtype=np.float32):
    """reference :27-66 -- the running acceptance ratio r_i and accepted mass p*_i of the first r_buffer_size steps,
    piecewise in closed form between consecutive sorted ratios.  NumPy; returns (r_buffer, pstar_buffer) of `dtype`."""
    log_ratios = np.asarray(torch.as_tensor(log_ratios).detach().cpu().numpy(), dtype=np.float32)
    t_mass = np.asarray(torch.as_tensor(t_mass).detach().cpu().numpy(), dtype=np.float64)
    p_mass = np.asarray(torch.as_tensor(p_mass).detach().cpu().numpy(), dtype=np.float64)
    ratios = np.exp(log_ratios)
    t_cum = np.exp(_cumulative_logsumexp(t_mass))
    p_cum_all = np.exp(_cumulative_logsumexp(p_mass))
    p_zero = float(1. - np.exp(np.logaddexp.reduce(p_mass)))
    pstar_buffer = np.zeros(r_buffer_size, dtype=dtype)
    r_buffer = np.zeros(r_buffer_size, dtype=dtype)
    r = 1.
    r_buffer[0] = r
    i = 1
    n = ratios.shape[0]
    for r_ind, r_next in enumerate(ratios):
        if r_next < r:
            continue
        p_cum = p_zero + (p_cum_all[r_ind - 1] if r_ind > 0 else 0.)
        t_c = t_cum[r_ind - 1] if r_ind > 0 else 0.
        fix = (1. - t_c) / (1. - p_cum)                       # fixed point of r <- p_cum * r + (1 - t_c)
        if r_ind == n - 1:
            if not math.isclose(r_next, fix, rel_tol=1e-5):
                raise CodingError("rejection sampler: last stratum is inconsistent with the accumulated masses")
            interval = r_buffer_size - i
        else:
            interval = min(r_buffer_size - i,
                           int(math.ceil(np.log((r_next - fix) / (r - fix)) // np.log(p_cum))))
        steps = 1. + np.arange(interval, dtype=dtype)
        r_slice = -np.exp(np.log(p_cum) * steps + np.log(fix - r)) + fix
        r_buffer[i:i + interval] = r_slice
        pstar_buffer[i - 1:i + interval - 1] = (1. - p_cum) * r_buffer[i - 1:i + interval - 1] + t_c
        r = np.power(p_cum, interval) * (r - fix) + fix
        i += interval
        if i == r_buffer_size:
            pstar_buffer[r_buffer_size - 1] = (1. - p_cum) * r + t_c
            break
        if r_ind == n - 1:
            raise CodingError('R Buffer incomplete after processing all samples. This is a bug.')
    return r_buffer, pstar_buffer


def gaussian_rejection_sample_small(t_dist, p_dist, sample_buffer_size, r_buffer_size, sample_generator, seed=42069):
    """reference :69-116 -> (index, sample).  O(e^KL) work: the caller partitions large Gaussians (GaussianCoder)."""
    assert r_buffer_size % sample_buffer_size == 0
    if tuple(t_dist.loc.shape) != tuple(p_dist.loc.shape):
        raise CodingError("target and proposal must have the same shape")
    log_ratios, t_mass, p_mass = get_t_p_mass(t_dist, p_dist, n_samples=100, oversampling=100, seed=seed)
    r_buffer, pstar_buffer = get_r_pstar(log_ratios, t_mass, p_mass, r_buffer_size=r_buffer_size)
    t_loc, t_scale = E._f32c(t_dist.loc, "cuda"), E._f32c(t_dist.scale, "cuda")
    p_loc, p_scale = E._f32c(p_dist.loc, "cuda"), E._f32c(p_dist.scale, "cuda")
    dl = torch.log(t_scale) - torch.log(p_scale)
    kl = float((0.5 * ((t_loc - p_loc) / p_scale) ** 2 + 0.5 * torch.expm1(2. * dl) - dl).sum())
    if kl >= 20.:
        raise CodingError('KL divergence={} is too high for rejection sampling'.format(kl))
    dev = t_loc.device
    r_dev = torch.from_numpy(np.asarray(r_buffer, np.float32)).to(dev)
    ps_dev = torch.from_numpy(np.asarray(pstar_buffer, np.float32)).to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed((int(seed) * 2654435761 + 1) & 0x7fffffffffffffff)
    i = 0
    for _ in range(int(r_buffer_size // sample_buffer_size)):
        sample_ratios = sample_generator.get_ratios(t_dist, p_dist, seed=seed + i // sample_buffer_size)
        accepted = (torch.exp(sample_ratios) - r_dev[i:i + sample_buffer_size]) / (1. - ps_dev[i:i + sample_buffer_size]) + \
            torch.rand(sample_ratios.shape, generator=gen, device=dev)
        hits = (accepted > 0.).nonzero()
        if hits.shape[0] > 0:
            index = int(hits[0, 0])
            return i + index, sample_generator.get_index(index)
        i += sample_buffer_size
    # beyond the buffer: accept anything above the last ratio
    log_r = math.log(float(r_buffer[-1]))
    while True:
        sample_ratios = sample_generator.get_ratios(t_dist, p_dist, seed=seed + i // sample_buffer_size)
        hits = (sample_ratios > log_r).nonzero()
        if hits.shape[0] > 0:
            index = int(hits[0, 0])
            return i + index, sample_generator.get_index(index)
        i += sample_buffer_size
