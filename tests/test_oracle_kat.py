"""CPU tests that PIN the oracle: reference known-answer vectors and self-consistency.

Golden material from the reference repository:
  * notebooks/Discrete REC.ipynb:51,64-66 -- 100 Bernoulli(0.7) draws after tf.random.set_seed(42)
    (pins Philox constants, key/counter layout, op-seed derivation, Uint32ToFloat, element<->counter map)
  * notebooks/scratch.ipynb:403-412 -- index alphabet {0..19} at Omega=3, eps=0 (S = int(e^3) = 20)
  * rec/coding/tests/test_coder.py:12-21 -- beam-search round trip (Omega=6, B=10, t=N(5.1,0.001), seed 69420)
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import oracle as O
import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

NOTEBOOK_BERNOULLI = [1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1,
                      1, 1, 1, 1, 1, 1, 0, 1, 0, 1, 0, 1, 1, 1, 0, 0, 1, 1, 1, 0, 0, 1,
                      0, 1, 1, 1, 0, 1, 1, 1, 0, 0, 1, 0, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1,
                      1, 1, 0, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 0, 1, 1, 0, 1, 1, 1, 1, 0,
                      1, 1, 1, 0, 1, 1, 1, 0, 0, 1, 1, 1]


def test_philox_random123_kat():
    assert [hex(x) for x in O.philox((0, 0), (0, 0, 0, 0))] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']


def test_notebook_bernoulli_vector():
    u = O.tf_uniform_f32_unseeded(42, 100)
    assert ((u < np.float32(0.7)).astype(int) == np.array(NOTEBOOK_BERNOULLI)).all()


def test_op_seed_is_python_randint():
    for s in (0, 1, 42, 43, 69420, 2 ** 31, 2 ** 40 + 5, 123456789012):
        assert O.py_randint31(s) == random.Random(s).randint(0, 2 ** 31 - 1)
    assert O.py_randint31(42) == 478163327


def test_scratch_notebook_alphabet():
    assert O.beam_num_samples(3., 1.) == 20
    assert O.beam_num_samples(3., 1.2) == 36


def test_ndtri_table_vs_scipy():
    from scipy.special import ndtri
    T = O.ndtri_table()
    k = np.arange(1, 10007)
    p32 = (k.astype(np.float32) / np.float32(10007)).astype(np.float64)
    ref = ndtri(p32)
    rel = np.abs(T[1:] - ref) / np.abs(ref)
    assert rel.max() < 1e-6
    assert abs(T[1] - (-3.7191932)) < 1e-6 and abs(T[10006] - 3.7191253) < 1e-6


def test_reference_unit_test_case_roundtrip():
    """rec/coding/tests/test_coder.py:12-21"""
    S = O.beam_num_samples(6., 1.)
    assert S == 403
    r = O.beam_encode_block([5.1], [0.001], [0.], [1.], 6., S, 10, 69420)
    assert r["n_aux"] == 4 and abs(r["kl"] - 19.41) < 0.01
    d = O.beam_decode_block([0.], [1.], S, 69420, r["indices"])
    assert np.array_equal(d.view(np.uint32), r["sample"].view(np.uint32))
    assert abs(float(d[0]) - 5.1) < 0.01


@pytest.mark.parametrize("D,B", [(1, 1), (64, 1), (64, 20), (100, 5), (1000, 3)])
def test_beam_roundtrip_bit_exact(D, B):
    tl, ts, pl, ps = synth.c2(D, data_seed=D + B)
    S = 36
    r = O.beam_encode_block(tl, ts, pl, ps, 3., S, B, 42)
    d = O.beam_decode_block(pl, ps, S, 42, r["indices"])
    assert np.array_equal(d.view(np.uint32), r["sample"].view(np.uint32))
    assert (r["indices"] >= 0).all() and (r["indices"] < S).all()


def test_is_roundtrip_bit_exact():
    mu, sig, pl, ps = synth.c1()
    S = O.is_num_samples(3 / np.log(2))
    r = O.is_encode_block(mu, sig, pl, ps, 3., S, 42)
    d = O.is_decode_block(pl, ps, 42, r["indices"])
    assert np.array_equal(d.view(np.uint32), r["sample"].view(np.uint32))


def test_canonical_score_matches_reference_form():
    """the centred quadratic (+ dropped constant) equals the reference's two-log_prob form: one candidate through the
    stand-alone reference-form evaluator, then ALL S*B candidates of EVERY partition of several coder-blocks against the
    exact (float64 throughout) log-ratio at north_star's 1e-5 relative tolerance (tests/test_refform_gap.py holds the
    population study)."""
    tl, ts, pl, ps = synth.c2(200, data_seed=5)
    r = O.beam_encode_block(tl, ts, pl, ps, 3., 36, 4, 7, trace=True)
    n_aux = r["n_aux"]
    sa, A, E, M = O.beam_schedule(tl, ts, pl, ps, n_aux)
    # final partition: the winning beam's value is sample - p_loc
    t = n_aux - 1
    x = (r["sample"] - pl).astype(np.float32)
    can = r["trace_score"][t, 0] + O.beam_score_constant(tl, ts, pl, ps, n_aux, t)
    ref = O.beam_refform_logw(tl, ts, pl, ps, n_aux, t, x)
    assert abs(can - ref) <= 1e-5 * max(1.0, abs(ref))
    for recipe, D, B, S, seed in (("c2", 200, 4, 36, 7), ("c2", 1000, 20, 36, 42), ("c3", 288, 10, 20, 7), ("c1", 64, 20, 36, 42)):
        tl, ts, pl, ps = getattr(synth, recipe)(D, data_seed=5)
        st = O.beam_refform_study(tl, ts, pl, ps, 3., S, B, seed)
        assert st["tf_partitions"] == st["n_aux"] > 0
        assert st["max_dev_canon_exact"] <= 1e-5, (recipe, D, st["max_dev_canon_exact"])


def test_hash_and_shuffle_helpers():
    assert O.simple_hash([]) == 1
    assert O.simple_hash([3]) == (3 * 69) % 10006 + 1
    assert O.simple_hash([3, 5]) == (3 * 69 + 5 * 70) % 10006 + 1
    p = O.shuffle_perm(1000, 42)
    assert sorted(p.tolist()) == list(range(1000))
    assert not np.array_equal(p, np.arange(1000))


def test_golden_fixture_matches_oracle():
    """tests/golden/beam_golden.json was produced by tests/golden/make_golden.py (oracle outputs);
    guards the oracle against accidental drift."""
    with open(os.path.join(GOLDEN, "beam_golden.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        tl, ts, pl, ps = getattr(synth, case["recipe"])(case["D"], data_seed=case["data_seed"])
        r = O.beam_encode_block(tl, ts, pl, ps, case["omega"], case["S"], case["B"], case["seed"])
        assert r["indices"].tolist() == case["indices"]
        assert r["sample"].view(np.uint32).tolist() == case["sample_bits"]
