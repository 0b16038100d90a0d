"""Reference-STRUCTURED NumPy port of the beam-search coder (test/bench infrastructure only).

Keeps the op structure of rec/coding/beam_search_coder.py: a materialised [S, B, D] candidate tensor
(`get_pseudo_random_sample`, :37-51), two `log_prob` passes and their difference (:82-84), a full argsort
over S*B (:86), decode regenerating all S samples to use one (:139-144).  It is what BASELINE.md section 2
calls the "reference-faithful mode" CPU baseline: the closest thing to the reference's TensorFlow-eager
path that can run here (TF/TFP are not installable).

Float semantics differ from the canonical oracle (irec_oracle.c) only in the log-weight rounding: this
file evaluates the reference's two-log_prob form with NumPy float32 reductions, the oracle the centred
quadratic in a fixed order.  tests/test_ref_numpy.py checks that the two agree on every index unless the
competing log-weights are within the stated 1e-5 relative tolerance, and that samples given equal indices
are bit-identical.
"""
import numpy as np

from . import oracle as O

PRIME = 10007
_T = None


def _table():
    global _T
    if _T is None:
        _T = O.ndtri_table()
    return _T


def _log_prob(x, loc, scale):
    """TFP 0.9 Normal._log_prob in float32"""
    half_log_2pi = np.float32(0.5 * np.log(2. * np.pi))
    d = x / scale - loc / scale
    return np.float32(-0.5) * (d * d) - (half_log_2pi + np.log(scale))


def _simple_hash(matrix):
    """beam_search_coder.py:33-35 (int32, wrapping)"""
    m = np.asarray(matrix, np.int32)
    w = np.arange(69, 69 + m.shape[1], dtype=np.int32)
    with np.errstate(over="ignore"):
        s = (m * w).sum(axis=1, dtype=np.int32)
    return np.mod(s, np.int32(PRIME - 1)) + np.int32(1)


def _pseudo_random_sample(scale, n_samples, index_matrix, seed):
    """beam_search_coder.py:37-51 -> [S, B, D] float32"""
    D = scale.shape[-1]
    r = O.beam_uniform_int(seed, 0, n_samples * D).reshape(n_samples, 1, D).astype(np.int64)
    h = _simple_hash(index_matrix).astype(np.int64).reshape(1, -1, 1)
    k = np.mod(r * h, PRIME)
    p = k.astype(np.float32) / np.float32(PRIME)
    del p                                   # quantile(p) is the table lookup T[k] (10006 distinct inputs)
    return _table()[k] * scale.reshape(1, 1, D)


def aux_ratio(i):
    return np.float32(np.power(i + 1., -0.7864636765648174))


def encode_block(t_loc, t_scale, p_loc, p_scale, kl_per_partition, n_samples, n_beams, seed, n_aux=None, trace=None):
    """BeamSearchCoder.encode_block (beam_search_coder.py:53-122).  Returns (indices, sample)."""
    tl, ts, pl, ps = (np.asarray(a, np.float32).reshape(-1) for a in (t_loc, t_scale, p_loc, p_scale))
    if n_aux is None:
        n_aux = O.n_aux(O.kl(tl, ts, pl, ps), kl_per_partition)
    cv, tv = ps * ps, ts * ts
    cum = np.zeros_like(cv)
    beams = None
    beam_indices = np.zeros((1, 0), np.int32)
    for it, i in enumerate(range(n_aux - 1, -1, -1)):
        v = aux_ratio(i) * (cv - cum)
        tot = v + cum
        m = (tl - pl) * tot / cv
        s2 = tv * (tot * tot) / (cv * cv) + tot * (cv - tot) / cv
        q_scale, p_cum_scale = np.sqrt(s2), np.sqrt(tot)
        samples = _pseudo_random_sample(np.sqrt(v), n_samples, beam_indices, seed + it)      # [S, B', D]
        combined = samples if beams is None else beams[None, :, :] + samples
        log_probs = (_log_prob(combined, m, q_scale) - _log_prob(combined, np.float32(0), p_cum_scale)).sum(axis=2)
        flat = log_probs.reshape(-1)
        order = np.argsort(-flat, kind="stable")                    # tf.argsort DESCENDING: ties -> lowest index
        n_cur = combined.shape[1]
        best = order[:n_beams]
        b_ind, s_ind = best % n_cur, best // n_cur
        if trace is not None:
            trace.append(dict(flat=flat.copy(), best=best.copy()))
        beams = combined[s_ind, b_ind]
        beam_indices = np.concatenate([beam_indices[b_ind, :it], s_ind[:, None].astype(np.int32)], axis=1)
        cum = cum + v
    return beam_indices[0].tolist(), beams[0] + pl


def decode_block(p_loc, p_scale, n_samples, seed, indices):
    """BeamSearchCoder.decode_block (beam_search_coder.py:124-148), regenerating all S samples per partition"""
    pl, ps = (np.asarray(a, np.float32).reshape(-1) for a in (p_loc, p_scale))
    n_aux = len(indices)
    cv = ps * ps
    cum = np.zeros_like(cv)
    sample = np.zeros_like(pl)
    for it, i in enumerate(range(n_aux - 1, -1, -1)):
        v = aux_ratio(i) * (cv - cum)
        aux = _pseudo_random_sample(np.sqrt(v), n_samples, np.asarray([indices[:it]], np.int32).reshape(1, it), seed + it)
        sample = sample + aux[indices[it], 0]
        cum = cum + v
    return sample + pl
