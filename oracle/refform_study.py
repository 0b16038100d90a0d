"""Population study: how often does the canonical log-weight of the oracle (and of the CUDA kernels, which are bit-identical
to it) select differently from the reference's own float32 two-log_prob form (rec/coding/beam_search_coder.py:79-106)?

TEST INFRASTRUCTURE (oracle/): run by tests/test_refform_gap.py on a small population, by bench.py's cpu legs on a bounded
sample, and stand-alone for the committed full-population record:

    python -m oracle.refform_study --blocks 2048 --out profiles/r2_refform_study.json

Population: C2 coder-blocks (resnet_vae latents [16,16,32], block_size 1000 -> D = 1000 and the last block D = 192,
n_beams 20, S = 36, Omega = 3) and C3 coder-blocks (large_level_2_vae recipe, D = 1000 / 288 / 56, n_beams 10, S = 20),
all from the seeded recipes of tests/synth.py split with the coder's own permutation.
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import oracle as O      # noqa: E402
import synth                        # noqa: E402

SEED = 42


def population(n_blocks):
    """(name, tl, ts, pl, ps, omega, S, B) for n_blocks coder-blocks: alternating C2 tensors (9 blocks each) and C3 level
    tensors (blocks of 1000 plus the short tail)"""
    out = []
    tensor = 0
    while len(out) < n_blocks:
        if tensor % 3 != 2:                                   # two C2 tensors, then one C3 tensor
            n, bs, recipe, omega, S, B = 8192, 1000, synth.c2, 3.0, 36, 20
        else:
            n, bs, recipe, omega, S, B = (12288 if tensor % 2 else 8 * 12 * 128 + 56), 1000, synth.c3, 3.0, 20, 10
        tl, ts, pl, ps = recipe(n, data_seed=50000 + tensor)
        perm = O.shuffle_perm(n, SEED)
        for b0 in range(0, n, bs):
            sel = perm[b0:min(n, b0 + bs)]
            out.append((f"{recipe.__name__}/t{tensor}/b{b0 // bs}", tl[sel], ts[sel], pl[sel], ps[sel], omega, S, B))
            if len(out) >= n_blocks:
                break
        tensor += 1
    return out


def study_block(job, sum_mode):
    name, tl, ts, pl, ps, omega, S, B = job
    r = O.beam_refform_study(tl, ts, pl, ps, omega, S, B, SEED, sum_mode=sum_mode)
    kl64 = O.kl(tl, ts, pl, ps)
    n64 = O.n_aux(kl64, omega)
    n32 = {m: O.n_aux(np.float32(O.kl_f32(tl, ts, pl, ps, m)), omega) for m in ("sequential", "pairwise")}
    r.update(name=name, D=int(tl.size), n_aux_f64=int(n64), n_aux_f32=n32)
    return r


def run(n_blocks, sum_mode="pairwise", threads=None):
    jobs = population(n_blocks)
    threads = threads or (os.cpu_count() or 1)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(lambda j: study_block(j, sum_mode), jobs))
    dt = time.perf_counter() - t0
    return summarise(res, sum_mode, dt, threads)


def summarise(res, sum_mode, dt, threads):
    nb = len(res)
    parts = sum(r["tf_partitions"] for r in res)
    idx_total = sum(r["n_aux"] for r in res)
    idx_equal = sum(int((r["indices_canonical"] == r["indices_refform"]).sum()) for r in res)
    set_mis = sum(r["tf_set_mismatch"] for r in res)
    worst = max(r["tf_max_rel_gap"] for r in res)
    return {
        "what": "canonical log-weight (oracle == CUDA kernels) vs the reference's float32 two-log_prob form, "
                "rec/coding/beam_search_coder.py:79-106",
        "sum_mode": sum_mode, "blocks": nb, "threads": threads, "seconds": dt,
        "block_dims": sorted(set(r["D"] for r in res)),
        "free_running": {
            "blocks_identical": sum(r["free_identical"] for r in res),
            "blocks_identical_pct": 100.0 * sum(r["free_identical"] for r in res) / nb,
            "index_match_pct_refform": 100.0 * idx_equal / max(idx_total, 1),
            "note": "after the first differing partition the two runs code different beams; later positions differ by construction",
        },
        "teacher_forced": {
            "partitions": parts,
            "kept_set_identical": parts - set_mis,
            "teacher_forced_match_pct": 100.0 * (parts - set_mis) / max(parts, 1),
            "kept_set_mismatches": set_mis,
            "order_only_mismatches": sum(r["tf_order_mismatch"] for r in res),
            "best_candidate_mismatches": sum(r["tf_best_mismatch"] for r in res),
            "worst_relative_gap_of_a_mismatch": worst,
            "all_mismatches_within_1e-5": bool(worst < 1e-5),
        },
        "score_deviation": {
            "denominator": "max(1, |exact log-ratio|) per candidate; constants fixed through the best candidate of the partition",
            "canonical_vs_exact_float64_max": max(r["max_dev_canon_exact"] for r in res),
            "reference_float32_form_vs_exact_float64_max": max(r["max_dev_ref32_exact"] for r in res),
            "canonical_vs_reference_form_float64_sum_max": max(r["max_rel_score_dev"] for r in res),
        },
        "kl_float32_vs_float64": {
            "blocks_n_aux_differs_sequential": sum(int(r["n_aux_f32"]["sequential"] != r["n_aux_f64"]) for r in res),
            "blocks_n_aux_differs_pairwise": sum(int(r["n_aux_f32"]["pairwise"] != r["n_aux_f64"]) for r in res),
            "note": "the oracle sums KL in float64 and rounds once; TFP sums float32 terms -- a last-ulp difference at a multiple of "
                    "Omega changes n_aux and with it every index of that block",
        },
    }


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=2048)
    ap.add_argument("--sum-mode", default="pairwise", choices=list(O.SUM_MODES))
    ap.add_argument("--threads", type=int, default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rep = run(a.blocks, a.sum_mode, a.threads)
    txt = json.dumps(rep, indent=1)
    print(txt)
    if a.out:
        with open(a.out, "w") as f:
            f.write(txt + "\n")
