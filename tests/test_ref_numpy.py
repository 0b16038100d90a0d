"""Cross-check of the two independent CPU restatements: the canonical C oracle (irec_oracle.c) and the
reference-structured NumPy port (oracle/ref_numpy.py, the reference's op structure and two-log_prob form).

Index agreement is exact except where the competing log-weights differ by < 1e-5 relative (north_star's
stated tolerance); given equal indices the decoded samples are bit-identical."""
import numpy as np
import pytest

import synth
from oracle import oracle as O
from oracle import ref_numpy as R


@pytest.mark.parametrize("recipe,D,B,S,omega,seed", [("c1", 64, 1, 36, 3., 42), ("c1", 64, 20, 36, 3., 42),
                                                     ("c2", 192, 20, 36, 3., 42), ("c3", 288, 10, 20, 3., 7),
                                                     ("c2", 1000, 20, 36, 3., 42),
                                                     ("c1", 64, 48, 36, 3., 42),       # n_beams > 32: the wide-beam kernel's range
                                                     ("c2", 100, 300, 7, 2., 5)])      # S < B: beams grow 1 -> 7 -> 49 -> 300
def test_numpy_port_agrees_with_oracle(recipe, D, B, S, omega, seed):
    tl, ts, pl, ps = getattr(synth, recipe)(D, data_seed=11)
    ref = O.beam_encode_block(tl, ts, pl, ps, omega, S, B, seed, trace=True)
    trace = []
    idx, sample = R.encode_block(tl, ts, pl, ps, omega, S, B, seed, trace=trace)
    if idx == ref["indices"].tolist():
        assert np.array_equal(sample.view(np.uint32), ref["sample"].view(np.uint32))
    else:
        # first partition where the kept sets differ must be a near tie in the NumPy log-weights
        for t, tr in enumerate(trace):
            nb = int(ref["trace_nbeams"][t])
            bcur = tr["flat"].size // S
            mine = set(int(f) for f in tr["best"][:nb])
            theirs = set(int(s) * bcur + int(b) for s, b in zip(ref["trace_s"][t, :nb], ref["trace_b"][t, :nb]))
            if mine != theirs:
                diff = sorted(mine ^ theirs)
                vals = tr["flat"][diff]
                scale = np.abs(tr["flat"][tr["best"][:nb]]).max()
                assert np.ptp(vals) <= 1e-5 * max(scale, 1.0), (t, diff, vals)
                break
    # decode of the NumPy port's own indices round-trips bit-exactly in both implementations
    d1 = R.decode_block(pl, ps, S, seed, idx)
    d2 = O.beam_decode_block(pl, ps, S, seed, idx)
    assert np.array_equal(d1.view(np.uint32), d2.view(np.uint32))
    assert np.array_equal(d1.view(np.uint32), sample.astype(np.float32).view(np.uint32))


def test_reference_unit_case_numpy():
    """rec/coding/tests/test_coder.py:12-21 through the reference-structured port"""
    idx, sample = R.encode_block([5.1], [0.001], [0.], [1.], 6., 403, 10, 69420)
    ref = O.beam_encode_block([5.1], [0.001], [0.], [1.], 6., 403, 10, 69420)
    assert idx == ref["indices"].tolist() == [369, 3, 318, 285]
    assert np.array_equal(sample.view(np.uint32), ref["sample"].view(np.uint32))
