"""CPU tests of the multi-GPU host logic (world_size 2, gloo, 127.0.0.1): range partitioning and the
record exchange of irec_b200/sharding.py -- the same helpers ShardedBeamBlock and bench.py use on NCCL.

The per-rank scoring is done by the CPU oracle (teacher-forced `beam_scores`) standing in for
irec_beam_step_score, and a NumPy merge standing in for k_topb_merge: what is under test is that a
candidate-range-sharded top-B + all-gather + merge reproduces the unsharded coder bit-for-bit (tie rule:
score descending, flat index s*B'+b ascending), through the product's buffer layout and collective call."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_ranges_cover_and_are_disjoint(built):
    from irec_b200 import sharding as SH
    for n in (0, 1, 7, 36, 1024, 2 ** 24 + 3):
        for world in (1, 2, 3, 4, 8):
            for fn in (SH.unit_range, SH.candidate_range):
                r = [fn(n, k, world) for k in range(world)]
                assert r[0][0] == 0 and r[-1][1] == n
                for a, b in zip(r, r[1:]):
                    assert a[1] == b[0] and a[0] <= a[1]
    assert SH.unit_range(10, 0, 4) == (0, 3) and SH.unit_range(10, 3, 4) == (8, 10)
    assert SH.candidate_range(36, 1, 8) == (5, 10) and SH.candidate_range(36, 7, 8) == (35, 36)


def _hash_from_sum(hsum):
    return int(np.int64(hsum) % 10006) + 1          # floor mod, like tf.math.floormod


def _sharded_oracle_encode(rank, world, dist, tl, ts, pl, ps, omega, S, B, seed):
    import torch
    from irec_b200 import sharding as SH
    from oracle import oracle as O
    T = O.ndtri_table()
    n_aux = O.n_aux(O.kl(tl, ts, pl, ps), omega)
    sa, A, E, M = O.beam_schedule(tl, ts, pl, ps, n_aux)
    D = tl.size
    beams, hsum, hist = None, np.zeros(1, np.int32), [[]]
    local, gathered = SH.alloc_exchange(B, world, "cpu")
    lo, hi = SH.candidate_range(S, rank, world)
    for t in range(n_aux):
        Bcur = hsum.size
        scores = O.beam_scores(T, sa[t], A[t], E[t], M[t], beams, hsum, S, seed + t)      # [S, Bcur]
        mine = [(-float(scores[s, b]), s * Bcur + b, s, b) for s in range(lo, hi) for b in range(Bcur)]
        mine.sort()
        mine = mine[:B]
        buf = np.zeros((B + 1, 4), np.int32)
        for i, (neg, f, s, b) in enumerate(mine):
            buf[i, 0] = np.float32(-neg).view(np.int32)
            buf[i, 1], buf[i, 2] = s, b
        buf[B, 0] = len(mine)
        local.copy_(torch.from_numpy(buf.reshape(-1)))
        recs, counts = SH.exchange_records(local, gathered, B, world, dist)
        recs = recs.numpy().reshape(world, B, 4)
        cand = []
        for r in range(world):
            for i in range(int(counts[r])):
                sc = recs[r, i, 0:1].view(np.float32)[0]
                cand.append((-float(sc), int(recs[r, i, 1]) * Bcur + int(recs[r, i, 2]), int(recs[r, i, 1]), int(recs[r, i, 2])))
        cand.sort()
        win = cand[:B]
        nb_, nh, nhist = [], [], []
        for _, _, s, b in win:
            r_ = O.beam_uniform_int(seed + t, s * D, D).astype(np.int64)
            k = (r_ * _hash_from_sum(hsum[b])) % 10007
            a = (T[k] * sa[t]).astype(np.float32)
            prev = np.zeros(D, np.float32) if beams is None else beams[b]
            nb_.append((prev + a).astype(np.float32))
            nh.append(np.int32((int(hsum[b]) + s * (69 + t) + 2 ** 31) % 2 ** 32 - 2 ** 31))
            nhist.append(hist[b] + [s])
        beams, hsum, hist = np.stack(nb_), np.asarray(nh, np.int32), nhist
    return hist[0], (beams[0] + pl).astype(np.float32)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "relative-entropy-coding_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = []
        for (recipe, D, S, B, omega, seed) in [("c1", 64, 36, 20, 3.0, 42), ("c2", 100, 7, 32, 4.0, 5), ("c1", 64, 403, 10, 6.0, 69420)]:
            tl, ts, pl, ps = getattr(synth, recipe)(D, data_seed=3)
            idx, sample = _sharded_oracle_encode(rank, world, dist, tl, ts, pl, ps, omega, S, B, seed)
            res.append((idx, sample))
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([(i, s.tobytes()) for i, s in res], dtype=object),
                allow_pickle=True)
    finally:
        dist.destroy_process_group()


def test_candidate_range_sharding_world2_gloo(built, tmp_path):
    import torch.multiprocessing as mp
    import synth
    from oracle import oracle as O
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(os.path.join(tmp_path, f"rank{r}.npy"), allow_pickle=True) for r in range(world)]
    cases = [("c1", 64, 36, 20, 3.0, 42), ("c2", 100, 7, 32, 4.0, 5), ("c1", 64, 403, 10, 6.0, 69420)]
    for ci, (recipe, D, S, B, omega, seed) in enumerate(cases):
        tl, ts, pl, ps = getattr(synth, recipe)(D, data_seed=3)
        ref = O.beam_encode_block(tl, ts, pl, ps, omega, S, B, seed)
        for r in range(world):
            idx, sample_bytes = outs[r][ci]
            assert list(idx) == ref["indices"].tolist(), (ci, r)
            assert sample_bytes == ref["sample"].tobytes(), (ci, r)
