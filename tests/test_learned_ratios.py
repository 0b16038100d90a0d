"""Learned auxiliary variance ratios (reference rec/coding/coder.py:197-410):
`GaussianCoder(extrapolate_auxiliary_ratios=False)` calibrates one variance ratio per auxiliary-variable index by SGD
and codes with that table instead of the power law.  CPU: the calibration loop and the error contract; the oracle with
a custom table.  GPU: kernels with the table == oracle with the table, bit for bit."""
import numpy as np
import pytest
import torch

import synth
from oracle import oracle as O
from rec.coding import BeamSearchCoder, GaussianCoder
from rec.coding.samplers import ImportanceSampler
from rec.coding.utils import CodingError
from irec_b200 import Normal


def _batch(n_blocks, D, seed, shift=0.6):
    rng = np.random.Generator(np.random.PCG64(seed))
    mu = (shift * rng.standard_normal((n_blocks, D))).astype(np.float32)
    sig = np.exp(rng.uniform(-1.2, 0, (n_blocks, D))).astype(np.float32)
    return mu, sig, np.zeros((n_blocks, D), np.float32), np.ones((n_blocks, D), np.float32)


def _fit(coder, seed=0, **kw):
    mu, sig, pl, ps = _batch(6, 64, seed)
    coder.update_block_auxiliary_variance_ratios(Normal(mu, sig), Normal(pl, ps), max_iters=400, learning_rate=0.01, **kw)
    return coder


def test_error_contract_and_fit_cpu():
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=5, extra_samples=1.2, extrapolate_auxiliary_ratios=False)
    with pytest.raises(CodingError):                     # reference :222-225
        coder.get_auxiliary_ratio(0)
    _fit(coder)
    r = coder.aux_variable_variance_ratios
    assert coder._initialized and r.dtype == np.float32 and r.shape[0] >= 4
    assert r[0] == 1.0 and np.all(r[1:] > 0.) and np.all(r[1:] < 1.)
    assert np.all(np.diff(r[1:]) < 0.05)                 # later (higher-index) variables take a smaller share
    with pytest.raises(CodingError):                     # reference :226-231
        coder.get_auxiliary_ratio(r.shape[0])
    counts = coder.average_counts.copy()
    _fit(coder, seed=1)                                  # running average over calls (reference :392-395)
    assert np.all(coder.average_counts[1:len(counts)] >= counts[1:]) and coder.average_counts[1] > counts[1]
    with pytest.raises(CodingError):
        BeamSearchCoder(kl_per_partition=3., n_beams=5).update_block_auxiliary_variance_ratios(None, None)


def test_fit_splits_the_kl_cpu():
    """the fitted ratio of the first coded auxiliary variable leaves ~Omega * (n - 1) nats for the rest"""
    coder = GaussianCoder(kl_per_partition=3., sampler=ImportanceSampler(coding_bits=5), extrapolate_auxiliary_ratios=False)
    mu, sig, pl, ps = _batch(8, 64, 3)
    coder.update_block_auxiliary_variance_ratios(Normal(mu, sig), Normal(pl, ps), max_iters=1500, learning_rate=0.02)
    from rec.coding.coder import get_auxiliary_coder, get_auxiliary_target
    t, c = Normal(mu, sig), Normal(pl, ps)
    kl = lambda q, p: (0.5 * ((q.loc - p.loc) / p.scale) ** 2 + 0.5 * torch.expm1(2 * (torch.log(q.scale) - torch.log(p.scale)))  # noqa: E731
                       - (torch.log(q.scale) - torch.log(p.scale))).sum(dim=1)
    total = kl(t, c)
    n = 1 + torch.floor(total / 3.).to(torch.int64)
    top = int(n.max())
    sel = n >= top
    v = float(coder.aux_variable_variance_ratios[top - 1]) * c.scale[sel] ** 2
    tt, cc = Normal(t.loc[sel], t.scale[sel]), Normal(c.loc[sel], c.scale[sel])
    aux_kl = kl(get_auxiliary_target(tt, cc, v), get_auxiliary_coder(cc, v))
    assert float(aux_kl.mean()) < 3. + 1.0 and float((total[sel] - aux_kl).mean()) < 3. * (top - 1) + 1.5


def test_oracle_with_learned_table_round_trips():
    mu, sig, pl, ps = synth.c1(64, data_seed=0)
    table = np.array([1.0, 0.52, 0.37, 0.3, 0.26, 0.22, 0.2, 0.18, 0.17, 0.16, 0.15, 0.14, 0.13, 0.12, 0.11, 0.1], np.float32)
    base = O.beam_encode_block(mu, sig, pl, ps, 3., 36, 5, 42)
    O.set_aux_ratios(table)
    try:
        enc = O.beam_encode_block(mu, sig, pl, ps, 3., 36, 5, 42)
        dec = O.beam_decode_block(pl, ps, 36, 42, enc["indices"])
    finally:
        O.set_aux_ratios(None)
    assert np.array_equal(dec.view(np.uint32), enc["sample"].view(np.uint32))
    assert enc["indices"].tolist() != base["indices"].tolist()          # the table really is in force
    again = O.beam_encode_block(mu, sig, pl, ps, 3., 36, 5, 42)
    assert again["indices"].tolist() == base["indices"].tolist()        # and really is cleared


@pytest.mark.gpu
def test_gpu_learned_table_matches_oracle(built):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = "cuda:0"
    mu, sig, pl, ps = _batch(5, 300, 11)
    for make in (lambda: BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2, extrapolate_auxiliary_ratios=False),
                 lambda: GaussianCoder(kl_per_partition=3., sampler=ImportanceSampler(coding_bits=6),
                                       extrapolate_auxiliary_ratios=False)):
        coder = make()
        t, c = Normal(mu, sig, device=dev), Normal(pl, ps, device=dev)
        coder.update_block_auxiliary_variance_ratios(t, c, max_iters=200, learning_rate=0.01)
        table = coder.aux_variable_variance_ratios.copy()
        beam = isinstance(coder, BeamSearchCoder)
        for b in range(3):
            tb, cb = Normal(mu[b:b + 1], sig[b:b + 1], device=dev), Normal(pl[b:b + 1], ps[b:b + 1], device=dev)
            indices, sample = coder.encode(tb, cb, seed=42)
            O.set_aux_ratios(table)
            try:
                if beam:
                    ref = O.beam_encode_block(mu[b], sig[b], pl[b], ps[b], 3., coder.n_samples, 20, 42)
                else:
                    ref = O.is_encode_block(mu[b], sig[b], pl[b], ps[b], 3., coder.sampler.n_samples, 42)
            finally:
                O.set_aux_ratios(None)
            assert [int(i) for i in indices] == [int(i) for i in ref["indices"]]
            assert np.array_equal(sample.cpu().numpy().reshape(-1).view(np.uint32), np.asarray(ref["sample"], np.float32).view(np.uint32))
            dec = coder.decode(cb, [int(i) for i in indices], seed=42)
            assert torch.equal(dec, sample)
        # the power-law coder gives a different index stream on the same input: the table is really used
        plain = BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2) if beam else \
            GaussianCoder(kl_per_partition=3., sampler=ImportanceSampler(coding_bits=6))
        i2, _ = plain.encode(Normal(mu[:1], sig[:1], device=dev), Normal(pl[:1], ps[:1], device=dev), seed=42)
        i1, _ = coder.encode(Normal(mu[:1], sig[:1], device=dev), Normal(pl[:1], ps[:1], device=dev), seed=42)
        assert [int(i) for i in i1] != [int(i) for i in i2]
        # a block that needs more auxiliary variables than the table holds: CodingError (reference :226-231)
        big = Normal(4. * mu[:1], sig[:1], device=dev)
        with pytest.raises(CodingError):
            coder.encode(big, Normal(pl[:1], ps[:1], device=dev), seed=42)
