#!/usr/bin/env python
"""bench_single.py -- single-image latency of BASELINE.json configs[1] and configs[2] through the callers' flow
(rec.models.LatentHierarchy: coder.encode per level, sequential levels, then .rec file; decode from the file).
Not the headline bench (bench.py, configs[3] batch throughput); one JSON line per config.

  configs[1]: resnet_vae, 24 latent tensors [16,16,32] (9 coder-blocks each), n_beams=20, extra_samples=1.2, Omega=3
  configs[2]: large_level_2_vae, [8,12,128] then [32,48,196] (13 + 302 coder-blocks), n_beams=10, extra_samples=1.0, Omega=3
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "relative-entropy-coding_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CONFIGS = {
    "configs[1] resnet_vae 32x32": dict(shapes=[(16, 16, 32)] * 24, recipe="c2", n_beams=20, extra=1.2, image=(32, 32, 3)),
    "configs[2] large_level_2_vae 768x512": dict(shapes=[(8, 12, 128), (32, 48, 196)], recipe="c3", n_beams=10, extra=1.0,
                                                 image=(512, 768, 3)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--max-aux", type=int, default=256, help="pipelined compress (no host sync between levels); 0 = level by level")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build()
    from irec_b200 import native as N
    from rec.coding import BeamSearchCoder
    from rec.models import LatentHierarchy, SyntheticLadder
    dev = "cuda:0"
    tmp = tempfile.mkdtemp()
    for name, c in CONFIGS.items():
        coder = BeamSearchCoder(kl_per_partition=3., n_beams=c["n_beams"], extra_samples=c["extra"], block_size=1000)
        model = LatentHierarchy(SyntheticLadder(c["shapes"], recipe=c["recipe"], data_seed=5, device=dev))
        path = os.path.join(tmp, "x.rec")
        enc, dec, launches = [], [], 0
        for rep in range(args.reps + 1):                       # first repetition is the warm-up
            torch.cuda.synchronize()
            l0 = N.launch_count()
            t0 = time.perf_counter()
            block_indices, latents = model.compress(42, coder, file_path=path, image_shape=c["image"], max_aux=args.max_aux or None)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            decoded = model.decompress(coder, file_path=path)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            launches = N.launch_count() - l0
            if rep:
                enc.append(t1 - t0)
                dec.append(t2 - t1)
        assert all(torch.equal(a, b) for a, b in zip(latents, decoded))
        n_aux = [len(b) for t in block_indices for b in t]
        S, B = coder.n_samples, coder.n_beams
        cand = sum(S + (k - 1) * S * min(B, S) for k in n_aux)
        dims = sum(int(np.prod(s)) for s in c["shapes"])
        print(json.dumps({"workload": name, "coder_blocks": len(n_aux), "partitions": int(sum(n_aux)), "candidates": int(cand),
                          "encode_ms": 1e3 * min(enc), "decode_ms": 1e3 * min(dec), "candidates_per_sec": cand / min(enc),
                          "partitions_per_sec": sum(n_aux) / min(enc), "file_bytes": os.path.getsize(path),
                          "bits_per_dim_file": 8 * os.path.getsize(path) / dims,
                          "bits_per_dim_index_stream": sum(n_aux) * np.log2(S) / dims, "gpu_launches": int(launches), "pipelined_levels": bool(args.max_aux),
                          "decode_bit_exact": True}), flush=True)


if __name__ == "__main__":
    main()
