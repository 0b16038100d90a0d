#!/usr/bin/env python
"""bench_sweep.py -- BASELINE.json configs[4]: large-partition sweep, ONE coder-block (D=64 synthetic posterior vs
N(0,I)), kl_per_partition from 3 to 20 bits (S = floor(exp(Omega_nats * 1.2)) up to ~1.7e7 candidates per auxiliary
variable), candidate index range split over the ranks with an NCCL all-gather of the per-rank top-B records
(irec_b200.engine.ShardedBeamBlock).  Not the headline bench (that is bench.py); prints one JSON line per Omega.

  python bench_sweep.py                       # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench_sweep.py
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=str, default="3,6,9,12,14,16,18,20")
    ap.add_argument("--beams", type=int, default=20)
    ap.add_argument("--dims", type=int, default=64)
    ap.add_argument("--min-aux", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true", help="compare with the CPU oracle where S is small enough")
    ap.add_argument("--graph", action="store_true", help="replay the per-variable step loop from a CUDA graph")
    ap.add_argument("--fused", action="store_true", help="one cooperative launch for all variables (k_gp_fused), as bench.py's c5")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import synth
    from irec_b200 import engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D, B = args.dims, args.beams
    for bits in [float(b) for b in args.bits.split(",")]:
        omega = np.float32(bits * np.log(2.0))
        S = int(np.exp(float(omega) * 1.2))
        mu, sig, pl, ps = synth.c1(D, data_seed=0)
        # scale the posterior mean so that the block needs >= min_aux auxiliary variables at this Omega
        kl0 = float(np.sum(0.5 * (mu.astype(np.float64) ** 2 + sig.astype(np.float64) ** 2 - 1) - np.log(sig.astype(np.float64))))
        need = args.min_aux * float(omega)
        if kl0 < need:
            base = float(np.sum(0.5 * (sig.astype(np.float64) ** 2 - 1) - np.log(sig.astype(np.float64))))
            scale = np.sqrt(max(need - base, 0.0) / max(float(np.sum(0.5 * mu.astype(np.float64) ** 2)), 1e-12)) * 1.02
            mu = (mu * scale).astype(np.float32)
        d = [torch.as_tensor(a, device=dev).contiguous() for a in (mu, sig, pl, ps)]
        blk = engine.ShardedBeamBlock(D, S, B, omega, max_aux=256, device=dev)
        idx, sample = blk.encode(*d, seed=42)            # warm-up + result
        if args.fused:
            idx_f, sample_f = blk.encode_fused(*d, seed=42)
            assert idx_f == idx and torch.equal(sample_f, sample), "fused encode differs from the multi-launch one"
        if args.graph:
            idx_g, sample_g = blk.encode_graphed(*d, seed=42)        # captures; must reproduce the eager result
            assert idx_g == idx and torch.equal(sample_g, sample), "graphed encode differs from the eager one"
        times = []
        for _ in range(args.reps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if args.fused:
                blk.encode_fused(*d, seed=42)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
                continue
            n_aux = blk.init(*d, seed=42)
            if args.graph:
                blk._graphs[n_aux].replay()
            else:
                for t in range(n_aux):
                    blk.step(t)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = torch.tensor([min(times)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms)
        n_aux = len(idx)
        cand = S + (n_aux - 1) * S * min(B, S)
        line = {"workload": "C5 sweep: one coder-block, candidate range sharded", "omega_bits": bits, "S": S, "D": D, "n_beams": B,
                "n_aux": n_aux, "n_gpus": world, "cuda_graph": bool(args.graph), "fused": bool(args.fused),
                "exchange": ("peer-memory stores (irec_p2p_exchange)" if blk.p2p is not None else ("nccl all_gather" if world > 1 else "none")), "ms": ms, "candidates_per_sec": cand / (ms * 1e-3),
                "candidate_dims_per_sec": cand * D / (ms * 1e-3), "partitions_per_sec": n_aux / (ms * 1e-3)}
        if args.check and S * B * D <= 4e8:
            from oracle import oracle as O
            ref = O.beam_encode_block(mu, sig, pl, ps, omega, S, B, 42, max_aux=256)
            line["matches_oracle"] = bool(idx == ref["indices"].tolist() and
                                          np.array_equal(sample.cpu().numpy().view(np.uint32), ref["sample"].view(np.uint32)))
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        blk._graphs.clear()
        del blk
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
