"""Run under torchrun on N GPUs: candidate-range-sharded beam encode (ShardedBeamBlock, NCCL all-gather of the
per-rank top-B records) must equal the CPU oracle bit-for-bit on every rank.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_sharded_nccl.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "relative-entropy-coding_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import synth
    from irec_b200 import engine
    from oracle import oracle as O
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for (D, S, B, omega_bits, seed) in [(64, 36, 20, None, 42), (64, 2000, 10, None, 42), (64, 50000, 20, None, 7), (300, 5000, 10, None, 3)]:
        omega = np.float32(np.log(S) / 1.2)
        mu, sig, pl, ps = synth.c1(D, data_seed=D + S)
        d = [torch.as_tensor(a, device=dev).contiguous() for a in (mu, sig, pl, ps)]
        blk = engine.ShardedBeamBlock(D, S, B, omega, max_aux=64, device=dev)
        idx, sample = blk.encode(*d, seed=seed)
        ref = O.beam_encode_block(mu, sig, pl, ps, omega, S, B, seed)
        same = idx == ref["indices"].tolist() and np.array_equal(sample.cpu().numpy().view(np.uint32), ref["sample"].view(np.uint32))
        # the one-launch path (irec_beam_encode_fused; peer-memory exchange inside the kernel) where it applies
        fused = "n/a"
        if blk.fused_available() and blk.world > 1:
            for rep in range(2):           # twice: the exchange buffers' sequence numbers carry over
                idx_f, sample_f = blk.encode_fused(*d, seed=seed)
                fused = bool(idx_f == idx and torch.equal(sample_f, sample))
                same = same and fused
        print(f"rank {rank}/{world} D={D} S={S} B={B} n_aux={len(idx)} range=[{blk.s_begin},{blk.s_end}) match={same} "
              f"fused={fused} exchange={'p2p' if blk.p2p is not None else 'nccl'}", flush=True)
        ok = ok and same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag) != 1:
        sys.exit(1)
    if rank == 0:
        print("sharded nccl ok")


if __name__ == "__main__":
    main()
