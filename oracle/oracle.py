"""ctypes front-end of the CPU oracle (oracle/irec_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package never imports this module.

Parity status: "parity unpinned" (see the header of irec_oracle.c and DESIGN.md): the reference's
TensorFlow code cannot run here; the RNG/seed plumbing is pinned by the reference's notebook vector.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

PRIME = 10007
CHUNK = 32

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.orc_py_randint31.restype = C.c_int64
        _lib.orc_py_randint31.argtypes = [C.c_int64]
        _lib.orc_ndtri_f32.restype = C.c_float
        _lib.orc_ndtri_f32.argtypes = [C.c_float]
        _lib.orc_aux_ratio.restype = C.c_float
        _lib.orc_aux_ratio.argtypes = [C.c_int]
        _lib.orc_kl.restype = C.c_float
        _lib.orc_n_aux.restype = C.c_int
        _lib.orc_n_aux.argtypes = [C.c_float, C.c_float]
        _lib.orc_simple_hash.restype = C.c_int32
        _lib.orc_is_num_samples.restype = C.c_int32
        _lib.orc_is_num_samples.argtypes = [C.c_float]
        _lib.orc_is_coded_sample.restype = C.c_int64
        _lib.orc_beam_refform_logw.restype = C.c_double
        _lib.orc_beam_score_constant.restype = C.c_double
        _lib.orc_kl_f32.restype = C.c_float
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


# ---------------------------------------------------------------------------------------------- RNG
def philox(key, ctr):
    out = np.zeros(4, np.uint32)
    c = np.asarray(ctr, np.uint32)
    lib().orc_philox_raw(C.c_uint32(key[0]), C.c_uint32(key[1]), _p(c, C.c_uint32), _p(out, C.c_uint32))
    return out


def tf_stream_u32(seed1, seed2, start, n):
    out = np.zeros(n, np.uint32)
    lib().orc_tf_stream_u32(C.c_int64(seed1), C.c_int64(seed2), C.c_int64(start), C.c_int64(n), _p(out, C.c_uint32))
    return out


def py_randint31(seed):
    return int(lib().orc_py_randint31(seed))


def tf_uniform_f32_unseeded(global_seed, n):
    out = np.zeros(n, np.float32)
    lib().orc_tf_uniform_f32_unseeded(C.c_int64(global_seed), C.c_int64(n), _p(out))
    return out


def beam_uniform_int(q, start, n):
    out = np.zeros(n, np.int32)
    lib().orc_beam_uniform_int(C.c_int64(q), C.c_int64(start), C.c_int64(n), _p(out, C.c_int32))
    return out


def is_normal_stream(seed, start, n):
    out = np.zeros(n, np.float32)
    lib().orc_is_normal_stream(C.c_int64(seed), C.c_int64(start), C.c_int64(n), _p(out))
    return out


def ndtri_table():
    T = np.zeros(PRIME, np.float32)
    lib().orc_ndtri_table(_p(T))
    return T


def ndtri_f32(p):
    return float(lib().orc_ndtri_f32(C.c_float(p)))


def aux_ratio(i):
    return float(lib().orc_aux_ratio(int(i)))


def set_aux_ratios(ratios=None):
    """learned auxiliary ratios for the calls that follow on this thread (None restores the power law)"""
    l = lib()
    l.orc_set_aux_ratios.restype = C.c_int
    l.orc_set_aux_ratios.argtypes = [C.c_void_p, C.c_int]
    if ratios is None:
        assert l.orc_set_aux_ratios(None, 0) == 0
        return
    r = _f32(ratios)
    assert l.orc_set_aux_ratios(_p(r), int(r.size)) == 0


def simple_hash(indices):
    a = np.ascontiguousarray(np.asarray(indices, np.int32))
    return int(lib().orc_simple_hash(_p(a, C.c_int32), C.c_int(a.size)))


def shuffle_perm(n, seed):
    out = np.zeros(n, np.int64)
    lib().orc_shuffle_perm(C.c_int64(n), C.c_int64(seed), _p(out, C.c_int64))
    return out


# ------------------------------------------------------------------------------------- schedule / KL
def kl(t_loc, t_scale, p_loc, p_scale):
    tl, ts, pl, ps = map(_f32, (t_loc, t_scale, p_loc, p_scale))
    return float(lib().orc_kl(_p(tl), _p(ts), _p(pl), _p(ps), C.c_int(tl.size)))


def n_aux(kl_value, omega):
    return int(lib().orc_n_aux(C.c_float(kl_value), C.c_float(omega)))


def beam_schedule(t_loc, t_scale, p_loc, p_scale, n_aux_):
    tl, ts, pl, ps = map(_f32, (t_loc, t_scale, p_loc, p_scale))
    D = tl.size
    sa = np.zeros((n_aux_, D), np.float32)
    A = np.zeros((n_aux_, D), np.float32)
    E = np.zeros((n_aux_, D), np.float32)
    M = np.zeros((n_aux_, D), np.float32)
    lib().orc_beam_schedule(_p(tl), _p(ts), _p(pl), _p(ps), C.c_int(D), C.c_int(n_aux_), _p(sa), _p(A), _p(E), _p(M))
    return sa, A, E, M


def beam_scores(T, sa_t, A_t, E_t, M_t, beams, hsum, S, q):
    """Teacher-forced scores of one partition.  beams: [B', D] or None (t = 0); hsum: int32 [B']."""
    sa_t, A_t, E_t, M_t = map(_f32, (sa_t, A_t, E_t, M_t))
    D = sa_t.size
    hs = np.ascontiguousarray(np.asarray(hsum, np.int32))
    Bcur = hs.size
    out = np.zeros((S, Bcur), np.float32)
    if beams is None:
        bp = C.POINTER(C.c_float)()
    else:
        beams = np.ascontiguousarray(np.asarray(beams, np.float32).reshape(Bcur, D))
        bp = _p(beams)
    lib().orc_beam_scores(_p(_f32(T)), _p(sa_t), _p(A_t), _p(E_t), _p(M_t), bp, _p(hs, C.c_int32), C.c_int(D),
                          C.c_int(S), C.c_int(Bcur), C.c_int64(q), _p(out))
    return out


def top_k_desc(v, k):
    v = _f32(v)
    out = np.zeros(k, np.int64)
    lib().orc_top_k_desc(_p(v), C.c_int64(v.size), C.c_int(k), _p(out, C.c_int64))
    return out


# ------------------------------------------------------------------------------------- beam coder
class OracleError(Exception):
    pass


class RefformStats(C.Structure):
    """orc_refform_stats_t (irec_oracle.c)"""
    _fields_ = [("n_aux", C.c_int32), ("free_identical", C.c_int32), ("free_first_diff", C.c_int32),
                ("tf_partitions", C.c_int32), ("tf_set_mismatch", C.c_int32), ("tf_order_mismatch", C.c_int32),
                ("tf_best_mismatch", C.c_int32), ("pad", C.c_int32), ("tf_max_rel_gap", C.c_double),
                ("max_rel_score_dev", C.c_double), ("max_dev_canon_exact", C.c_double), ("max_dev_ref32_exact", C.c_double)]


SUM_MODES = {"sequential": 0, "pairwise": 1, "float64": 2}


def beam_refform_study(t_loc, t_scale, p_loc, p_scale, omega, S, B, seed, sum_mode="pairwise", max_aux=4096):
    """canonical coder vs the reference's two-log_prob float32 form on one coder-block, free-running and teacher-forced
    (orc_beam_refform_study).  Returns a dict of the statistics plus both index lists."""
    tl, ts, pl, ps = map(_f32, (t_loc, t_scale, p_loc, p_scale))
    ic = np.zeros(max_aux, np.int32)
    ir = np.zeros(max_aux, np.int32)
    st = RefformStats()
    rc = lib().orc_beam_refform_study(_p(tl), _p(ts), _p(pl), _p(ps), C.c_int(tl.size), C.c_float(omega), C.c_int(S), C.c_int(B),
                                      C.c_int64(seed), C.c_int(SUM_MODES[sum_mode]), _p(ic, C.c_int32), _p(ir, C.c_int32),
                                      C.c_int(max_aux), C.byref(st))
    if rc != 0:
        raise OracleError(f"orc_beam_refform_study rc={rc} (n_aux={st.n_aux})")
    out = {name: getattr(st, name) for name, _ in RefformStats._fields_ if name != "pad"}
    out["indices_canonical"] = ic[:st.n_aux].copy()
    out["indices_refform"] = ir[:st.n_aux].copy()
    return out


def kl_f32(t_loc, t_scale, p_loc, p_scale, sum_mode="pairwise"):
    """KL(target || coder) the way TFP evaluates it in float32 (irec_oracle.c orc_kl_f32)"""
    tl, ts, pl, ps = map(_f32, (t_loc, t_scale, p_loc, p_scale))
    return float(lib().orc_kl_f32(_p(tl), _p(ts), _p(pl), _p(ps), C.c_int(tl.size), C.c_int(SUM_MODES[sum_mode])))


def beam_num_samples(kl_per_partition, extra_samples):
    """beam_search_coder.py:29  int(np.exp(kl_per_partition * extra_samples))"""
    return int(np.exp(kl_per_partition * extra_samples))


def beam_encode_block(t_loc, t_scale, p_loc, p_scale, omega, S, B, seed, max_aux=4096, trace=False):
    tl, ts, pl, ps = map(_f32, (t_loc, t_scale, p_loc, p_scale))
    D = tl.size
    idx = np.zeros(max_aux, np.int32)
    n = C.c_int32(0)
    klv = C.c_float(0)
    sample = np.zeros(D, np.float32)
    if trace:
        tsc = np.zeros((max_aux, B), np.float32)
        tss = np.zeros((max_aux, B), np.int32)
        tsb = np.zeros((max_aux, B), np.int32)
        tnb = np.zeros(max_aux, np.int32)
        targs = (_p(tsc), _p(tss, C.c_int32), _p(tsb, C.c_int32), _p(tnb, C.c_int32))
    else:
        targs = (C.POINTER(C.c_float)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)())
    rc = lib().orc_beam_encode(_p(tl), _p(ts), _p(pl), _p(ps), C.c_int(D), C.c_float(omega), C.c_int(S), C.c_int(B),
                               C.c_int64(seed), _p(idx, C.c_int32), C.c_int(max_aux), C.byref(n), _p(sample),
                               C.byref(klv), *targs)
    if rc != 0:
        raise OracleError(f"orc_beam_encode rc={rc} (n_aux={n.value}, kl={klv.value})")
    res = {"indices": idx[:n.value].copy(), "sample": sample, "kl": klv.value, "n_aux": n.value}
    if trace:
        res.update(trace_score=tsc[:n.value], trace_s=tss[:n.value], trace_b=tsb[:n.value], trace_nbeams=tnb[:n.value])
    return res


def beam_decode_block(p_loc, p_scale, S, seed, indices):
    pl, ps = map(_f32, (p_loc, p_scale))
    D = pl.size
    idx = np.ascontiguousarray(np.asarray(indices, np.int32))
    out = np.zeros(D, np.float32)
    lib().orc_beam_decode(_p(pl), _p(ps), C.c_int(D), C.c_int(S), C.c_int64(seed), _p(idx, C.c_int32),
                          C.c_int(idx.size), _p(out))
    return out


def beam_refform_logw(t_loc, t_scale, p_loc, p_scale, n_aux_, t, x):
    tl, ts, pl, ps, x = map(_f32, (t_loc, t_scale, p_loc, p_scale, x))
    return float(lib().orc_beam_refform_logw(_p(tl), _p(ts), _p(pl), _p(ps), C.c_int(tl.size), C.c_int(n_aux_),
                                             C.c_int(t), _p(x)))


def beam_score_constant(t_loc, t_scale, p_loc, p_scale, n_aux_, t):
    tl, ts, pl, ps = map(_f32, (t_loc, t_scale, p_loc, p_scale))
    return float(lib().orc_beam_score_constant(_p(tl), _p(ts), _p(pl), _p(ps), C.c_int(tl.size), C.c_int(n_aux_),
                                               C.c_int(t)))


# ------------------------------------------------------------------------------- importance sampler
def is_num_samples(coding_bits):
    return int(lib().orc_is_num_samples(C.c_float(coding_bits)))


def is_coded_sample(t_loc, t_scale, p_loc, p_scale, S, seed, want_scores=False):
    tl, ts, pl, ps = map(_f32, (t_loc, t_scale, p_loc, p_scale))
    D = tl.size
    out = np.zeros(D, np.float32)
    sc = np.zeros(S, np.float32) if want_scores else None
    idx = lib().orc_is_coded_sample(_p(tl), _p(ts), _p(pl), _p(ps), C.c_int(D), C.c_int32(S), C.c_int64(seed), _p(out),
                                    _p(sc) if want_scores else C.POINTER(C.c_float)())
    return (int(idx), out, sc) if want_scores else (int(idx), out)


def is_decode_sample(p_loc, p_scale, index, seed):
    pl, ps = map(_f32, (p_loc, p_scale))
    out = np.zeros(pl.size, np.float32)
    lib().orc_is_decode_sample(_p(pl), _p(ps), C.c_int(pl.size), C.c_int64(index), C.c_int64(seed), _p(out))
    return out


def is_encode_block(t_loc, t_scale, p_loc, p_scale, omega, S, seed, max_aux=4096):
    tl, ts, pl, ps = map(_f32, (t_loc, t_scale, p_loc, p_scale))
    D = tl.size
    idx = np.zeros(max_aux, np.int64)
    n = C.c_int32(0)
    klv = C.c_float(0)
    sample = np.zeros(D, np.float32)
    rc = lib().orc_is_encode_block(_p(tl), _p(ts), _p(pl), _p(ps), C.c_int(D), C.c_float(omega), C.c_int32(S),
                                   C.c_int64(seed), _p(idx, C.c_int64), C.c_int(max_aux), C.byref(n), _p(sample),
                                   C.byref(klv))
    if rc != 0:
        raise OracleError(f"orc_is_encode_block rc={rc}")
    return {"indices": idx[:n.value].copy(), "sample": sample, "kl": klv.value}


def is_decode_block(p_loc, p_scale, seed, indices):
    pl, ps = map(_f32, (p_loc, p_scale))
    idx = np.ascontiguousarray(np.asarray(indices, np.int64))
    out = np.zeros(pl.size, np.float32)
    lib().orc_is_decode_block(_p(pl), _p(ps), C.c_int(pl.size), C.c_int64(seed), _p(idx, C.c_int64), C.c_int(idx.size),
                              _p(out))
    return out
