#!/usr/bin/env python
"""bench_is.py -- the importance-sampling row of the hot path (SURVEY.md 8a a15/a16, 8d W_IS): GaussianCoder with an
ImportanceSampler (alpha = inf) over a batch of coder-blocks, encode and decode.  Not the headline bench (bench.py).

Workload: `--images` latent tensors [16,16,32] (C2 recipe), block_size 1000, kl_per_partition 3 nats, coding_bits so that
S = ceil(2^bits) matches the beam coder's budget (3 * 1.2 nats -> 5.2 bits -> S = 37).  One launch of `k_is_block` codes
every block: per auxiliary variable Philox4x32-10 + Box-Muller candidates (tf.random.normal restated), the canonical
log-weight, arg-max, on-device conditioning (coder.py:533-540).  Prints one JSON line:
candidates/s, candidate-dims/s, fraction of the INT32/FP32 issue roofline with W_IS = 34 lane-instr per candidate-dim,
decode time, and the C oracle on a bounded sample of the same blocks.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "relative-entropy-coding_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

LATENT, BLOCK, OMEGA, SEED = 8192, 1000, 3.0, 42
W_IS = 34.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--bits", type=float, default=3.0 * 1.2 / np.log(2.0))
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-blocks", type=int, default=64)
    args = ap.parse_args()

    import torch
    import __graft_entry__ as g
    g.build()
    import synth
    from irec_b200 import engine as E, native as N
    from rec.coding import GaussianCoder
    from rec.coding.samplers import ImportanceSampler
    from oracle import oracle as O

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    coder = GaussianCoder(kl_per_partition=OMEGA, sampler=ImportanceSampler(coding_bits=args.bits), block_size=BLOCK)
    S = coder.sampler.n_samples
    n_img = args.images
    arrs = [synth.c2(LATENT, data_seed=7000 + i) for i in range(n_img)]
    tl, ts, pl, ps = (torch.from_numpy(np.stack([a[k] for a in arrs])).to(dev).reshape(-1) for k in range(4))
    perm = coder._permutation(LATENT, SEED, dev)
    gather = (perm[None, :] + torch.arange(n_img, device=dev)[:, None] * LATENT).reshape(-1).contiguous()
    offsets, nb, max_dim = E.make_block_offsets(LATENT, BLOCK, dev, n_items=n_img)
    dims = (offsets[1:] - offsets[:-1]).cpu().numpy()

    def encode():
        return E.is_encode_blocks(tl, ts, pl, ps, gather, offsets, nb, max_dim, coder.kl_per_partition, S, SEED)

    indices, sample = encode()                      # warm-up + results
    n_aux = np.array([len(i) for i in indices])
    cand = int((n_aux * S).sum())
    cd = int((n_aux * S * dims).sum())
    torch.cuda.synchronize()
    l0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.reps):
        encode()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.reps
    launches = (N.launch_count() - l0) // args.reps
    # device time of the whole call: plan + candidate table + coding kernel, one packed D2H of the index rows
    ms_call = e0.elapsed_time(e1) / args.reps

    dec = E.is_decode_blocks(pl, ps, gather, offsets, nb, max_dim, SEED, indices)
    assert torch.equal(dec, sample), "decode != encode sample"
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        E.is_decode_blocks(pl, ps, gather, offsets, nb, max_dim, SEED, indices)
    torch.cuda.synchronize()
    dec_s = (time.perf_counter() - t0) / args.reps

    # CPU oracle on a bounded sample of the same blocks (and parity on them)
    hp = perm.cpu().numpy()
    tlh, tsh, plh, psh = (t.cpu().numpy() for t in (tl, ts, pl, ps))
    smp = sample.cpu().numpy()
    nblk_img = nb // n_img
    t0 = time.perf_counter()
    c_cand = 0
    match = 0
    for b in range(min(args.cpu_blocks, nb)):
        img, k = divmod(b, nblk_img)
        sel = img * LATENT + hp[k * BLOCK:min(LATENT, (k + 1) * BLOCK)]
        ref = O.is_encode_block(tlh[sel], tsh[sel], plh[sel], psh[sel], OMEGA, S, SEED)
        c_cand += len(ref["indices"]) * S
        ok = [int(i) for i in ref["indices"]] == [int(i) for i in indices[b]] and \
            np.array_equal(np.asarray(ref["sample"], np.float32).view(np.uint32), smp[sel].view(np.uint32))
        match += int(ok)
    cpu_s = time.perf_counter() - t0

    props = torch.cuda.get_device_properties(dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = props.multi_processor_count * 128 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    sec = ms_call * 1e-3
    print(json.dumps({
        "workload": f"importance sampler: {n_img} x [16,16,32] latents, block_size {BLOCK}, kl_per_partition {OMEGA}, "
                    f"coding_bits {args.bits:.3f} (S={S})",
        "coder_blocks": nb, "partitions": int(n_aux.sum()), "candidates": cand,
        "encode_ms_per_call": ms_call, "encode_wall_ms_per_call": 1e3 * wall, "gpu_launches_per_call": int(launches),
        "candidates_per_sec": cand / sec, "candidate_dims_per_sec": cd / sec, "partitions_per_sec": float(n_aux.sum()) / sec,
        "roofline": {"bound": "issue", "work_model": "W_IS = 34 lane-instr per candidate-dim (SURVEY.md 8d)",
                     "achieved_g_lane_instr_s": cd * W_IS / sec / 1e9, "peak_g_lane_instr_s": peak / 1e9,
                     "frac": cd * W_IS / sec / peak},
        "decode_ms_per_call": 1e3 * dec_s, "decode_bit_exact": True,
        "cpu_baseline": {"kind": "port", "cores": 1, "sample": f"{min(args.cpu_blocks, nb)} coder-blocks, C oracle, 1 thread",
                         "candidates_per_sec": c_cand / cpu_s, "blocks_identical_to_gpu": f"{match}/{min(args.cpu_blocks, nb)}"},
    }))


if __name__ == "__main__":
    main()
