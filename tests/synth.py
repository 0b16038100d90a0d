"""Seeded synthetic posteriors/priors of SURVEY.md section 8(d) (numpy PCG64, float32)."""
import numpy as np


def c1(D=64, data_seed=0):
    """C1/C5: p = N(0, I), mu_t ~ 0.6 N(0,1), sigma_t = exp(U(-1.2, 0))"""
    rng = np.random.Generator(np.random.PCG64(data_seed))
    mu = (0.6 * rng.standard_normal(D)).astype(np.float32)
    sig = np.exp(rng.uniform(-1.2, 0, D)).astype(np.float32)
    return mu, sig, np.zeros(D, np.float32), np.ones(D, np.float32)


def c2(n=8192, data_seed=0):
    """C2/C4 latent tensor: prior mu_p ~ 0.5 N, sigma_p = exp(0.3 N); posterior mu_t = mu_p + sigma_p 0.4 N,
    sigma_t = sigma_p exp(U(-1, 0))"""
    rng = np.random.Generator(np.random.PCG64(data_seed))
    pl = (0.5 * rng.standard_normal(n)).astype(np.float32)
    ps = np.exp(0.3 * rng.standard_normal(n)).astype(np.float32)
    tl = (pl + ps * 0.4 * rng.standard_normal(n)).astype(np.float32)
    ts = (ps * np.exp(rng.uniform(-1.0, 0.0, n))).astype(np.float32)
    return tl, ts, pl, ps


def c3(n, data_seed=0):
    """C3: same recipe with sigma = softplus(N(0,1)) + 1e-7 (large_2_level_vae.py:333,349,364,380)"""
    rng = np.random.Generator(np.random.PCG64(data_seed))
    sp = lambda x: np.log1p(np.exp(x))
    pl = (0.5 * rng.standard_normal(n)).astype(np.float32)
    ps = (sp(rng.standard_normal(n)) + 1e-7).astype(np.float32)
    tl = (pl + ps * 0.4 * rng.standard_normal(n)).astype(np.float32)
    ts = (ps * np.exp(rng.uniform(-1.0, 0.0, n))).astype(np.float32)
    return tl, ts, pl, ps
