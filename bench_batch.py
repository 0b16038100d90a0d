#!/usr/bin/env python
"""bench_batch.py -- BASELINE.json configs[3] as a pipeline with REAL level-to-level dependence: a batch of images through a
24-level ladder whose priors are computed from the previous level's coded latent (rec.models.BatchedSyntheticLadder, the
stand-in for the RVAE's residual blocks, resnet_vae.py:803-836), coded by LatentHierarchy.compress_batch: one launch per level
and sub-batch, the sub-batches pipelined on CUDA streams, no host synchronisation between levels.  Not the headline bench
(bench.py times the same launches with pre-resident inputs); prints one JSON line per stream count."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=128)
    ap.add_argument("--levels", type=int, default=24)
    ap.add_argument("--streams", type=str, default="1,2")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build()
    from rec.coding import BeamSearchCoder
    from rec.models import BatchedSyntheticLadder, LatentHierarchy
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2, block_size=1000)
    model = LatentHierarchy(BatchedSyntheticLadder([(16, 16, 32)] * args.levels, args.images, recipe="c2", data_seed=0, device=dev))
    ref = None
    for k in [int(x) for x in args.streams.split(",")]:
        idx, _ = model.compress_batch(seed=42, coder=coder, n_streams=k)          # warm-up + result
        if ref is None:
            ref = idx
        same = idx == ref
        n_aux = np.array([len(b) for img in idx for lvl in img for b in lvl], np.int64)
        cand = int((coder.n_samples + (n_aux - 1) * coder.n_samples * min(coder.n_beams, coder.n_samples)).sum())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            model.compress_batch(seed=42, coder=coder, n_streams=k)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.reps
        print(json.dumps({"workload": f"compress_batch: {args.images} images x {args.levels} dependent levels [16,16,32], block_size 1000, "
                                      f"n_beams 20, S {coder.n_samples}", "n_streams": k, "ms": 1e3 * dt, "candidates_per_sec": cand / dt,
                          "partitions_per_sec": int(n_aux.sum()) / dt, "identical_to_first": bool(same),
                          "includes": "ladder kernels, index read-back and Python list building"}), flush=True)


if __name__ == "__main__":
    main()
