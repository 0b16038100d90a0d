#!/usr/bin/env python
"""bench.py -- throughput of the iREC encode hot path on B200 (BASELINE.json metric).

Headline workload = BASELINE.json configs[3] itself: a batch of 1024 ImageNet32-shaped synthetic images through resnet_vae
latents -- 24 latent tensors [16,16,32] per image, block_size 1000 (9 coder-blocks per tensor), beam search n_beams=20,
extra_samples=1.2 (S=36), kl_per_partition=3 nats, coding seed 42 -- sharded over the N ranks by images (strong scaling:
1024 images in total whatever N; no collective in the loop).  One "step" = one pass of the encoder over the whole batch: 24
sequential level launches (a level's prior depends on the previous level's sample in the real model), every launch coding
(1024/N) x 9 coder-blocks.

  value    : candidates scored / s, inputs resident in HBM (CUDA events, max over ranks)
  e2e      : same through the public API (rec.coding.BeamSearchCoder.encode_batch) from pinned host buffers,
             host<->device copies and index read-back inside the timed region, over all --steps
  roofline : INT32/FP32 issue rate of the dominant kernel, algorithmic instructions per candidate-dim
             W = 10 + 24/B' (SURVEY.md 8d); secondary: HBM bytes and the shared-memory wavefront rate (profiles/)
  c5       : BASELINE.json configs[4] -- ONE coder-block (D=64), kl_per_partition 14..20 bits (S up to 2^24 candidates per
             auxiliary variable), the candidate index range split over the N ranks, per-rank top-B records exchanged over
             NVLink peer memory (irec_p2p_exchange) and merged; oracle check where S is small, 1-GPU equality at S = 2^24
  is       : the importance-sampler coder (GaussianCoder + ImportanceSampler) on 256 latents, encode + decode
  cpu_baseline / cpu_baselines, refform : CPU legs on rank 0 at N=1 (C oracle on all host threads; the reference-structured
             NumPy port; canonical-vs-reference-form decision statistics on a bounded sample)

`--impl reference` times the CPU port of the reference (the C oracle, all host threads) on a bounded sample of the same
workload and prints the same JSON line with "impl": "reference".
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "relative-entropy-coding_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

IMAGES_TOTAL = 1024
LEVELS = 24
LATENT = 8192            # 16*16*32
BLOCK = 1000
OMEGA, EXTRA, NBEAMS, SEED = 3.0, 1.2, 20, 42
S = int(np.exp(OMEGA * EXTRA))
MAX_AUX = 256
METRIC = "iREC candidates scored/sec (beam-search encode; candidate = one (sample, beam) pair scored over all dims of its coder-block)"
C5_BITS = (14.0, 16.0, 18.0, 20.0)
C5_DIMS, C5_MIN_AUX = 64, 8
IS_IMAGES = 256
W_IS = 34.0              # lane-instructions per candidate-dim of the importance sampler


def synth_level(image, level, n=LATENT):
    """SURVEY.md 8(d) C2 recipe, one latent tensor"""
    import synth
    return synth.c2(n, data_seed=1000 * image + level)


def synth_levels(images, threads):
    """host[level] = list over images of (tl, ts, pl, ps); numpy releases the GIL in the generators"""
    from concurrent.futures import ThreadPoolExecutor
    jobs = [(i, lvl) for lvl in range(LEVELS) for i in images]
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        res = list(ex.map(lambda j: synth_level(*j), jobs, chunksize=64))
    n = len(images)
    return [res[lvl * n:(lvl + 1) * n] for lvl in range(LEVELS)]


def work_model(n_aux, dims, s=S, nbeams=NBEAMS):
    """algorithmic work of a set of blocks: candidates, candidate-dims, partitions, lane-instructions"""
    n_aux = np.asarray(n_aux, np.int64)
    dims = np.asarray(dims, np.int64)
    bp = min(nbeams, s)
    cand = s * 1 + (n_aux - 1) * s * bp                     # t = 0 has one (empty) beam
    cd = cand * dims
    instr = dims * (s * 1 * (10 + 24 / 1) + (n_aux - 1) * s * bp * (10 + 24 / bp))
    return int(cand.sum()), int(cd.sum()), int(n_aux.sum()), float(instr.sum())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU legs (oracle/)
def cpu_sample_jobs(n_blocks):
    """`n_blocks` coder-blocks of the headline workload (full 1000-dim blocks of images 0, 1, ...)"""
    from oracle import oracle as O
    perm = O.shuffle_perm(LATENT, SEED)
    jobs = []
    for i in range(n_blocks):
        tl, ts, pl, ps = synth_level(i // 8, i % LEVELS)
        sel = perm[(i % 8) * BLOCK:(i % 8 + 1) * BLOCK]
        jobs.append((tl[sel], ts[sel], pl[sel], ps[sel]))
    return jobs


def cpu_port_sample(n_blocks, threads):
    """oracle (C port of the reference) on `n_blocks` coder-blocks of the same workload, `threads` host threads"""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.lib()
    O.ndtri_table()
    jobs = cpu_sample_jobs(n_blocks)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(lambda j: O.beam_encode_block(*j, OMEGA, S, NBEAMS, SEED), jobs))
    dt = time.perf_counter() - t0
    cand, cd, parts, _ = work_model([r["n_aux"] for r in res], [BLOCK] * n_blocks)
    return cand / dt, cd / dt, parts / dt, dt, res, jobs


def cpu_numpy_port_sample(n_blocks, threads):
    """the reference-STRUCTURED NumPy port (oracle/ref_numpy.py: materialised [S,B,D] candidates, two log_prob passes, full
    argsort -- the op structure of rec/coding/beam_search_coder.py:53-122), block-parallel on `threads` host threads"""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ref_numpy as R
    jobs = cpu_sample_jobs(n_blocks)
    R.encode_block(*jobs[0], OMEGA, S, NBEAMS, SEED, n_aux=2)          # warm the tables
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(lambda j: R.encode_block(*j, OMEGA, S, NBEAMS, SEED), jobs))
    dt = time.perf_counter() - t0
    cand, cd, parts, _ = work_model([len(r[0]) for r in res], [BLOCK] * n_blocks)
    return cand / dt, cd / dt, dt


def cpu_sample_blocks(cores):
    """bounded CPU sample: ~0.4 s of single-core work per coder-block -> 10-20 s per step on all cores"""
    return int(min(1024, max(32 * cores, 64)))


def auto_streams(n_img, sm_count=148):
    """sub-batches per rank: a launch of fewer than ~8 coder-blocks per block context (two contexts per SM) is split in two,
    so that the CTAs of the next launch take over the SMs the current one's tail leaves idle.  Measured on one B200
    (profiles/r2_streams_sweep.json): 128 images per rank (1152 blocks, 3.9 per context) 2.69e9 -> 3.04e9 candidates/s
    with 2 streams (3: 3.00e9, 4: 2.97e9); 256 and 1024 images per rank: no difference (3.05e9 either way)."""
    return 2 if 9 * n_img <= 8 * 2 * sm_count else 1


def workload_name():
    return (f"BASELINE configs[3]: batch of {IMAGES_TOTAL} ImageNet32-shaped synthetic images x {LEVELS} resnet_vae latents "
            f"[16,16,32], block_size {BLOCK}, beam_search n_beams={NBEAMS} extra_samples={EXTRA} (S={S}) kl_per_partition={OMEGA}")


def config_dict(world):
    """identical for the b200 arm and the reference arm of the same N"""
    return {"workload": workload_name(), "images_total": IMAGES_TOTAL,
            "parallelism": f"dp{world} (images sharded over the ranks, no collective in the loop)",
            "l2": "inputs of one step (3.2 GB / N per GPU) exceed the 126 MB L2"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_blocks = cpu_sample_blocks(cores)
    vals = []
    for _ in range(1 if args.warmup > 0 else 0):
        cpu_port_sample(n_blocks, cores)
    t_all = 0.0
    for _ in range(args.steps):
        v, cd, parts, dt, _, _ = cpu_port_sample(n_blocks, cores)
        vals.append((v, cd, parts))
        t_all += dt
    v = float(np.mean([x[0] for x in vals]))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "candidates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(int(os.environ.get("WORLD_SIZE", "1"))),
        "candidate_dims_per_sec": float(np.mean([x[1] for x in vals])),
        "partitions_per_sec": float(np.mean([x[2] for x in vals])),
        "cpu_baseline": {"value": v, "unit": "candidates/s", "cores": cores, "kind": "port",
                         "sample": f"{n_blocks} coder-blocks (D=1000, S=36, B=20) of the workload per step, C oracle "
                                   f"(oracle/irec_oracle.c), {cores} threads"},
        "e2e": {"value": v, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "TensorFlow 2.1/TFP 0.9 are not installable here; this is the C port of the reference algorithm "
                "(oracle/irec_oracle.c) run block-parallel on all host threads -- a stronger CPU baseline than the "
                "reference-structured NumPy port (bench.py's cpu_baselines[1])",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ C5: candidate-range sharding
def c5_posterior(bits):
    """SURVEY.md 8(d) C5 recipe: D=64 posterior vs N(0,I), mean scaled so that the block needs >= C5_MIN_AUX variables"""
    import synth
    omega = np.float32(bits * np.log(2.0))
    mu, sig, pl, ps = synth.c1(C5_DIMS, data_seed=0)
    mu64, sig64 = mu.astype(np.float64), sig.astype(np.float64)
    kl0 = float(np.sum(0.5 * (mu64 ** 2 + sig64 ** 2 - 1) - np.log(sig64)))
    need = C5_MIN_AUX * float(omega)
    if kl0 < need:
        base = float(np.sum(0.5 * (sig64 ** 2 - 1) - np.log(sig64)))
        scale = np.sqrt(max(need - base, 0.0) / max(float(np.sum(0.5 * mu64 ** 2)), 1e-12)) * 1.02
        mu = (mu * scale).astype(np.float32)
    return omega, mu, sig, pl, ps


def bench_c5(ctx, args):
    """configs[4] under the driver's clock: per Omega one coder-block, every rank scores a contiguous candidate range, the
    per-rank top-B records are exchanged (peer-memory stores, else NCCL all-gather) and merged identically on every rank;
    the per-variable loop is replayed from a CUDA graph.  Times: CUDA events, mean over reps, max over ranks."""
    import torch
    import torch.distributed as dist
    from irec_b200 import engine
    dev, rank, world = ctx["dev"], ctx["rank"], ctx["world"]
    out = []
    for bits in C5_BITS:
        omega, mu, sig, pl, ps = c5_posterior(bits)
        S5 = int(np.exp(float(omega) * EXTRA))
        d = [torch.as_tensor(a, device=dev).contiguous() for a in (mu, sig, pl, ps)]
        blk = engine.ShardedBeamBlock(C5_DIMS, S5, NBEAMS, omega, max_aux=256, device=dev)
        idx, sample = blk.encode(*d, seed=SEED)                   # multi-launch state machine, eager: warm-up + result
        fused = blk.fused_available()
        if fused:
            idx_g, sample_g = blk.encode_fused(*d, seed=SEED)     # one cooperative launch for all variables
        else:
            idx_g, sample_g = blk.encode_graphed(*d, seed=SEED)   # captures the loop; must reproduce the eager result
        graph_ok = bool(idx_g == idx and torch.equal(sample_g, sample))
        times = []
        for _ in range(args.c5_reps):
            ctx["barrier"]()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if fused:
                blk.encode_fused(*d, seed=SEED)
            else:
                n_aux = blk.init(*d, seed=SEED)
                blk._graphs[n_aux].replay()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms)
        n_aux = len(idx)
        cand = S5 + (n_aux - 1) * S5 * min(NBEAMS, S5)
        rec = {"omega_bits": bits, "S": S5, "D": C5_DIMS, "n_beams": NBEAMS, "n_aux": n_aux, "ms": ms,
               "candidates_per_sec": cand / (ms * 1e-3), "candidate_dims_per_sec": cand * C5_DIMS / (ms * 1e-3),
               "exchange": ("peer-memory stores (irec_p2p_exchange)" if blk.p2p is not None else
                            ("nccl all_gather" if world > 1 else "none")),
               "launch": ("one cooperative launch for all variables (k_gp_fused)" if fused else "7 launches per variable, CUDA graph"),
               "equals_multi_launch_path": graph_ok}
        if rank == 0:
            if S5 <= 1.2e5:
                from oracle import oracle as O
                ref = O.beam_encode_block(mu, sig, pl, ps, omega, S5, NBEAMS, SEED, max_aux=256)
                rec["matches_oracle"] = bool(idx == ref["indices"].tolist() and
                                             np.array_equal(sample.cpu().numpy().view(np.uint32), ref["sample"].view(np.uint32)))
            if world > 1 and bits == C5_BITS[-1]:
                one = engine.ShardedBeamBlock(C5_DIMS, S5, NBEAMS, omega, max_aux=256, device=dev, single=True)
                idx1, sample1 = one.encode(*d, seed=SEED)
                rec["equals_1gpu_result"] = bool(idx1 == idx and torch.equal(sample1, sample))
                del one
        torch.cuda.synchronize()
        blk._graphs.clear()
        del blk
        ctx["barrier"]()
        out.append(rec)
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    peak = world * sm_count * 128 * ctx["sm_max_mhz"] * 1e6
    top = out[-1]
    w = 10 + 24 / NBEAMS
    return {"workload": "BASELINE configs[4]: one coder-block (D=64 posterior vs N(0,I)), candidate index range split over the "
                        "ranks, top-B records exchanged and merged per auxiliary variable",
            "n_gpus": world, "sweep": out,
            "roofline": {"bound": "issue", "kernel": "k_gp_fused<20>" if "k_gp_fused" in top["launch"] else "k_gp_score_topb2<20>",
                         "at_omega_bits": top["omega_bits"],
                         "achieved": top["candidate_dims_per_sec"] * w / 1e9, "peak": peak / 1e9, "unit": "G lane-instr/s",
                         "frac": top["candidate_dims_per_sec"] * w / peak,
                         "work_model": "W = 10 + 24/B' lane-instr per candidate-dim, peak = N x SMs x 128 lanes x sm_max_mhz"}}


# ------------------------------------------------------------------------------------------------ importance sampler
def bench_is(ctx, args):
    """GaussianCoder + ImportanceSampler (rec/coding/coder.py:493-584, importance_sampling.py:9-103): IS_IMAGES latents
    [16,16,32] in total (sharded by images), block_size 1000, S = ceil(2^coding_bits) = 37; encode (plan + candidate table +
    k_is_block, one packed D2H) and decode (k_is_decode), CUDA events, max over ranks"""
    import torch
    import torch.distributed as dist
    import synth
    from irec_b200 import engine as E, native as N
    from irec_b200.sharding import unit_range
    from rec.coding import GaussianCoder
    from rec.coding.samplers import ImportanceSampler
    dev, rank, world = ctx["dev"], ctx["rank"], ctx["world"]
    lo, hi = unit_range(IS_IMAGES, rank, world)
    n_img = hi - lo
    coder = GaussianCoder(kl_per_partition=OMEGA, sampler=ImportanceSampler(coding_bits=OMEGA * EXTRA / np.log(2.0)), block_size=BLOCK)
    S_is = coder.sampler.n_samples
    arrs = [synth.c2(LATENT, data_seed=7000 + i) for i in range(lo, hi)]
    tl, ts, pl, ps = (torch.from_numpy(np.stack([a[k] for a in arrs])).to(dev).reshape(-1) for k in range(4))
    perm = coder._permutation(LATENT, SEED, dev)
    gather = (perm[None, :] + torch.arange(n_img, device=dev)[:, None] * LATENT).reshape(-1).contiguous()
    offsets, nb, max_dim = E.make_block_offsets(LATENT, BLOCK, dev, n_items=n_img)
    dims = (offsets[1:] - offsets[:-1]).cpu().numpy()

    def encode():
        return E.is_encode_blocks(tl, ts, pl, ps, gather, offsets, nb, max_dim, coder.kl_per_partition, S_is, SEED)

    indices, sample = encode()                      # warm-up + results (also settles the row-capacity hint)
    n_aux = np.array([len(i) for i in indices])
    cand, cd = int((n_aux * S_is).sum()), int((n_aux * S_is * dims).sum())
    ctx["barrier"]()
    l0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.is_reps):
        encode()
    e1.record()
    torch.cuda.synchronize()
    launches = (N.launch_count() - l0) // args.is_reps
    enc_ms = e0.elapsed_time(e1) / args.is_reps
    dec = E.is_decode_blocks(pl, ps, gather, offsets, nb, max_dim, SEED, indices)
    decode_ok = bool(torch.equal(dec, sample))
    packed = E.pack_indices(indices, torch.int64, dev)
    ws = torch.empty(int(N.lib().irec_is_block_workspace_bytes(packed[2])), dtype=torch.uint8, device=dev)
    out = torch.empty_like(pl)
    ctx["barrier"]()
    t0 = time.perf_counter()
    for _ in range(args.is_reps):
        E.is_decode_blocks(pl, ps, gather, offsets, nb, max_dim, SEED, indices)
    torch.cuda.synchronize()
    dec_wall_ms = 1e3 * (time.perf_counter() - t0) / args.is_reps
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for _ in range(args.is_reps):            # the launch alone, index rows already on the device
        N.check(N.lib().irec_is_decode(N.ptr(pl), N.ptr(ps), N.ptr(gather), N.ptr(offsets), nb, int(max_dim), SEED, N.ptr(packed[0]),
                                       packed[2], N.ptr(packed[1]), N.ptr(out), None, N.ptr(ws), ws.numel(), N.stream_ptr()), "irec_is_decode")
    d1.record()
    torch.cuda.synchronize()
    dec_kernel_ms = d0.elapsed_time(d1) / args.is_reps
    decode_ok = decode_ok and bool(torch.equal(out, sample))
    t = torch.tensor([enc_ms, dec_wall_ms, dec_kernel_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([cand, cd, float(n_aux.sum()), nb], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    enc_ms, dec_wall_ms, dec_kernel_ms = (float(x) for x in t)
    cand_all, cd_all, parts_all, nb_all = (float(x) for x in tot)
    sec = enc_ms * 1e-3
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    peak = world * sm_count * 128 * ctx["sm_max_mhz"] * 1e6
    res = {"workload": f"importance sampler: {IS_IMAGES} x [16,16,32] latents in total, block_size {BLOCK}, kl_per_partition {OMEGA}, "
                       f"coding_bits {OMEGA * EXTRA / np.log(2.0):.3f} (S={S_is})",
           "n_gpus": world, "coder_blocks": int(nb_all), "partitions": int(parts_all), "candidates": int(cand_all),
           "encode_ms_per_call": enc_ms, "gpu_launches_per_call": int(launches),
           "candidates_per_sec": cand_all / sec, "candidate_dims_per_sec": cd_all / sec, "partitions_per_sec": parts_all / sec,
           "roofline": {"bound": "issue", "kernel": "k_is_block", "work_model": "W_IS = 34 lane-instr per candidate-dim (SURVEY.md 8d): "
                        "Philox + Box-Muller + score per candidate-dim as the reference evaluates them; the kernel reads the launch-wide "
                        "candidate table instead (4 B per candidate-dim from L2)",
                        "achieved": cd_all * W_IS / sec / 1e9, "peak": peak / 1e9, "unit": "G lane-instr/s", "frac": cd_all * W_IS / sec / peak,
                        "l2_read_gbs": cd_all * 4 / sec / 1e9 / world},
           "decode_ms_per_call": dec_wall_ms, "decode_kernel_ms": dec_kernel_ms, "decode_bit_exact": decode_ok}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from concurrent.futures import ThreadPoolExecutor
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        n_cpu = min(nb, max(64, 4 * cores))
        hp = perm.cpu().numpy()
        tlh, tsh, plh, psh = (x.cpu().numpy() for x in (tl, ts, pl, ps))
        smp = sample.cpu().numpy()
        nblk_img = nb // n_img

        def one(b):
            img, k = divmod(b, nblk_img)
            sel = img * LATENT + hp[k * BLOCK:min(LATENT, (k + 1) * BLOCK)]
            ref = O.is_encode_block(tlh[sel], tsh[sel], plh[sel], psh[sel], OMEGA, S_is, SEED)
            ok = [int(i) for i in ref["indices"]] == [int(i) for i in indices[b]] and \
                np.array_equal(np.asarray(ref["sample"], np.float32).view(np.uint32), smp[sel].view(np.uint32))
            return len(ref["indices"]) * S_is, int(ok)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            rr = list(ex.map(one, range(n_cpu)))
        dt = time.perf_counter() - t0
        res["blocks_identical_to_oracle"] = f"{sum(r[1] for r in rr)}/{n_cpu}"
        res["index_match_pct"] = 100.0 * sum(r[1] for r in rr) / n_cpu
        res["cpu_baseline"] = {"value": sum(r[0] for r in rr) / dt, "unit": "candidates/s", "cores": cores, "kind": "port",
                               "sample": f"{n_cpu} coder-blocks of the same workload, C oracle, {cores} threads, {dt:.1f} s"}
    return res


# ------------------------------------------------------------------------------------------------ main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--images-total", type=int, default=IMAGES_TOTAL, help="configs[3] is 1024; smaller values are for kernel-tuning runs")
    ap.add_argument("--streams", type=int, default=0, help="sub-batches per rank, each on its own stream (0 = by launch size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="kernel-tuning runs: skip the end-to-end leg")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--no-is", action="store_true")
    ap.add_argument("--c5-reps", type=int, default=3)
    ap.add_argument("--is-reps", type=int, default=5)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    from irec_b200 import engine as E, native as N, Normal
    from irec_b200.sharding import unit_range
    from rec.coding import BeamSearchCoder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    images_total = args.images_total
    img_lo, img_hi = unit_range(images_total, rank, world)
    n_img = img_hi - img_lo
    assert n_img > 0, "more ranks than images"
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ctx = {"dev": dev, "rank": rank, "world": world, "barrier": barrier, "sm_max_mhz": sm_max_mhz}

    # ---------------- synthetic inputs (pinned host + device copies) ----------------
    cores = os.cpu_count() or 1
    levels = synth_levels(list(range(img_lo, img_hi)), max(1, cores // max(1, min(world, 8))))
    host = []      # per level: 4 pinned tensors [n_img, LATENT]
    for lvl in range(LEVELS):
        arrs = levels[lvl]
        host.append([torch.from_numpy(np.stack([a[k] for a in arrs])).pin_memory() for k in range(4)])
    del levels
    devt = [[t.to(dev, non_blocking=True) for t in lv] for lv in host]
    coder = BeamSearchCoder(kl_per_partition=OMEGA, n_beams=NBEAMS, extra_samples=EXTRA, block_size=BLOCK)
    perm = coder._permutation(LATENT, SEED, dev)
    gather = (perm[None, :] + torch.arange(n_img, device=dev)[:, None] * LATENT).reshape(-1).contiguous()
    offsets, nb, max_dim = E.make_block_offsets(LATENT, BLOCK, dev, n_items=n_img)
    dims = (offsets[1:] - offsets[:-1]).cpu().numpy()
    lib = N.lib()
    out_idx = [torch.empty((nb, MAX_AUX), dtype=torch.int32, device=dev) for _ in range(LEVELS)]
    out_na = [torch.empty(nb, dtype=torch.int32, device=dev) for _ in range(LEVELS)]
    out_st = [torch.empty(nb, dtype=torch.int32, device=dev) for _ in range(LEVELS)]
    out_sample = [torch.empty(n_img * LATENT, dtype=torch.float32, device=dev) for _ in range(LEVELS)]
    path = int(lib.irec_beam_encode_path(nb, max_dim, S, NBEAMS))
    kernel_name = {2: "k_beam_encode_resident2<20>", 3: "k_beam_encode_tmem<20>", 1: "k_beam_encode_resident<20>"}.get(path, f"path {path}")

    # The images of a rank are coded as `n_sub` independent sub-batches, each on its own CUDA stream: the levels of a
    # sub-batch stay sequential (a level's prior needs the previous level's sample), but sub-batch k's level l + 1 does not
    # wait for sub-batch j's level l.  A launch is a persistent grid of one CTA per SM whose CTAs retire when the block queue
    # is empty, so the next stream's CTAs take over the SMs one by one: the tail of one launch (a rank's 1152 coder-blocks
    # are 3.9 per block context at N = 8) is filled by the head of the next.
    n_sub = max(1, min(args.streams if args.streams > 0 else auto_streams(n_img), n_img))
    blocks_per_image = nb // n_img
    subs = []
    for k in range(n_sub):
        lo, hi = unit_range(n_img, k, n_sub)
        offs_k, nb_k, max_dim_k = E.make_block_offsets(LATENT, BLOCK, dev, n_items=hi - lo)
        ws_k = torch.empty(int(lib.irec_beam_encode_workspace_bytes(nb_k, max_dim_k, S, NBEAMS, MAX_AUX)), dtype=torch.uint8, device=dev)
        subs.append({"lo": lo, "hi": hi, "nb": nb_k, "max_dim": max_dim_k, "offsets": offs_k, "ws": ws_k,
                     "gather": gather[:(hi - lo) * LATENT], "b0": lo * blocks_per_image,
                     "stream": torch.cuda.Stream(device=dev) if n_sub > 1 else None})

    def launch_level(lvl, sub):
        lo, hi, b0, nbk = sub["lo"], sub["hi"], sub["b0"], sub["nb"]
        tl, ts, pl, ps = (t[lo:hi].reshape(-1) for t in devt[lvl])
        stream = sub["stream"].cuda_stream if sub["stream"] is not None else N.stream_ptr()
        N.check(lib.irec_beam_encode(N.ptr(tl), N.ptr(ts), N.ptr(pl), N.ptr(ps), N.ptr(sub["gather"]), N.ptr(sub["offsets"]), nbk,
                                     sub["max_dim"], OMEGA, S, NBEAMS, SEED, N.ptr(out_idx[lvl][b0:b0 + nbk]), MAX_AUX,
                                     N.ptr(out_na[lvl][b0:b0 + nbk]), N.ptr(out_st[lvl][b0:b0 + nbk]),
                                     N.ptr(out_sample[lvl][lo * LATENT:hi * LATENT]), N.ptr(sub["ws"]), sub["ws"].numel(), stream),
                "irec_beam_encode")

    def step_resident():
        if n_sub == 1:
            for lvl in range(LEVELS):
                launch_level(lvl, subs[0])
            return
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        for sub in subs:
            sub["stream"].wait_event(fork)
        for lvl in range(LEVELS):
            for sub in subs:
                launch_level(lvl, sub)
        for sub in subs:
            join = torch.cuda.Event()
            join.record(sub["stream"])
            cur.wait_event(join)

    # ---------------- device-resident throughput ----------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = N.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        step_resident()
    stop.record()
    barrier()
    launches = N.launch_count() - launches0
    ms_total = start.elapsed_time(stop)
    clock_info = clocks.stop() if rank == 0 else None
    st_all = torch.stack(out_st).cpu().numpy()
    na_all = torch.stack(out_na).cpu().numpy()
    assert (st_all == 0).all(), "some coder-blocks failed (status != 0)"
    cand, cd, parts, instr = 0, 0, 0, 0.0
    per_level_instr = []
    for lvl in range(LEVELS):
        c, d_, p_, i_ = work_model(na_all[lvl], dims)
        cand += c; cd += d_; parts += p_; instr += i_
        per_level_instr.append(i_)
    if os.environ.get("IREC_BENCH_DUMP") and rank == 0:      # per-level work of this run (used by profiles/make_bench_launch_json.py)
        per_level = [work_model(na_all[lvl], dims) for lvl in range(LEVELS)]
        json.dump({"images_per_gpu": n_img, "levels": [{"candidates": c, "candidate_dims": d_, "partitions": p_} for c, d_, p_, _ in per_level]},
                  open(os.environ["IREC_BENCH_DUMP"], "w"))
    # The timed region holds nothing but the level launches, back to back (and overlapping where n_sub > 1): the dominant
    # kernel's time per LEVEL (all sub-batch launches of a level) is the CUDA-event time of the region / (steps x levels).
    kernel_ms_avg = ms_total / (args.steps * LEVELS)
    achieved_instr = float(np.mean(per_level_instr)) / (kernel_ms_avg * 1e-3)

    # ---------------- end-to-end through the public API (host buffers) ----------------
    sample_host = [torch.empty((n_img, LATENT), dtype=torch.float32, pin_memory=True) for _ in range(LEVELS)]

    def step_e2e():
        """all levels through the public API from pinned host buffers; the index lists of level l are read back and built
        while level l + 1 runs (encode_batch(lazy=True)); every level's inputs cross H2D and its sample and indices D2H.
        With n_sub > 1 every sub-batch runs this pipeline on its own stream (torch.cuda.stream), as in step_resident."""
        total, pending = 0, []
        for lvl in range(LEVELS):
            now = []
            for sub in subs:
                lo, hi = sub["lo"], sub["hi"]
                with torch.cuda.stream(sub["stream"]) if sub["stream"] is not None else contextlib.nullcontext():
                    tl, ts, pl, ps = (t[lo:hi].to(dev, non_blocking=True) for t in host[lvl])
                    get_indices, sample = coder.encode_batch(Normal(tl, ts), Normal(pl, ps), seed=SEED, lazy=True)
                    sample_host[lvl][lo:hi].copy_(sample, non_blocking=True)
                now.append(get_indices)
            for get in pending:
                total += sum(len(b) for img in get() for b in img)
            pending = now
        for get in pending:
            total += sum(len(b) for img in get() for b in img)
        torch.cuda.synchronize()
        return total, sample_host[-1]

    e2e_s = float("nan")
    if not args.no_e2e:
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            parts_e2e, _ = step_e2e()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        assert parts_e2e == parts, (parts_e2e, parts)
        for lvl in (0, LEVELS - 1):
            assert torch.equal(sample_host[lvl].reshape(-1), out_sample[lvl].cpu()), "e2e sample differs from the resident run"
    h2d = LEVELS * 4 * n_img * LATENT * 4
    d2h = LEVELS * (n_img * LATENT * 4 + nb * (2 + int(E._hint_get((str(dev), float(OMEGA))))) * 4)

    # ---------------- max over ranks, sums over ranks ----------------
    t_res = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([cand, cd, parts, h2d, d2h, launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms = float(t_res[0]), float(t_res[1])
    cand_all, cd_all, parts_all, h2d_all, d2h_all, launches_all = (float(x) for x in tot)
    sec = ms_total * 1e-3 / args.steps

    # free the headline workload before the sub-benches
    del devt, out_idx, out_sample, host, sample_host
    torch.cuda.empty_cache()
    c5 = None if args.no_c5 else bench_c5(ctx, args)
    isb = None if args.no_is else bench_is(ctx, args)

    if rank == 0:
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        peak_instr = sm_count * 128 * sm_max_mhz * 1e6
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes_level = float(16 * dims.sum() + 4 * dims.sum() + 4 * na_all.mean(axis=0).sum())
        # measured per-launch DRAM traffic and shared-memory wavefronts per candidate-dim of the same kernel, from the
        # committed `ncu --set full` capture (profiles/<round>_bench_launch_ncu.json)
        traffic, traffic_src, wf_per_cd = None, None, None
        for name in ("r2_bench_launch_ncu.json", "r1_bench_launch_ncu.json"):
            try:
                prof = json.load(open(os.path.join(ROOT, "profiles", name)))
                if prof.get("kernel", "k_beam_encode_resident2<20>") != kernel_name:
                    continue
                if prof.get("images_per_gpu") == n_img:
                    traffic = float(prof["dram_bytes_read"] + prof["dram_bytes_write"])
                else:       # other launch size: scale the measured bytes per candidate-dim
                    traffic = float(prof["dram_bytes_read"] + prof["dram_bytes_write"]) / float(prof["candidate_dims"]) * (cd / LEVELS)
                wf_per_cd = float(prof["shared_wavefronts"]) / float(prof["candidate_dims"])
                traffic_src = prof.get("source")
                break
            except Exception:
                continue
        smem_rate = (cd_all / world / sec) * wf_per_cd if wf_per_cd else float("nan")
        line = {
            "metric": METRIC, "value": cand_all / sec, "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config_dict(world), images_total=images_total),
            "candidate_dims_per_sec": cd_all / sec, "partitions_per_sec": parts_all / sec,
            "index_match_pct": None,
            "gpu_launches": int(launches_all),
            "build_mode": g.BUILD_MODE,
            "clocks": clock_info,
            "e2e": {"value": cand_all / (e2e_ms * 1e-3), "unit": "candidates/s", "h2d_bytes_per_step": int(h2d_all),
                    "d2h_bytes_per_step": int(d2h_all), "ms_per_step": e2e_ms, "steps": args.steps},
            "roofline": {"bound": "issue", "kernel": kernel_name, "achieved": achieved_instr / 1e9,
                         "peak": peak_instr / 1e9, "unit": "G lane-instr/s", "frac": achieved_instr / peak_instr,
                         "peak_source": f"{sm_count} SMs x 128 lanes x {sm_max_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz)",
                         "work_model": "W = 10 + 24/B' lane-instr per candidate-dim (SURVEY.md 8d)",
                         "avg_launch_ms": kernel_ms_avg, "avg_launch_ms_is": "time of the timed region / (steps x levels): all launches of one level",
                         "blocks_per_launch": int(nb) // n_sub, "launches_per_level": n_sub, "traffic": traffic, "traffic_source": traffic_src,
                         "smem": {"wavefronts_per_candidate_dim": wf_per_cd, "achieved_gwf_s": smem_rate / 1e9,
                                  "peak_gwf_s": sm_count * sm_max_mhz * 1e6 / 1e9,
                                  "frac": smem_rate / (sm_count * sm_max_mhz * 1e6), "source": traffic_src},
                         "hbm": {"achieved_gbs": alg_bytes_level / (kernel_ms_avg * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "frac": alg_bytes_level / (kernel_ms_avg * 1e-3) / 1e9 / hbm_peak,
                                 "algorithmic_bytes_per_launch": alg_bytes_level}},
        }
        if c5 is not None:
            line["c5"] = c5
        if isb is not None:
            line["is"] = isb
        if world == 1 and not args.no_cpu_baseline:
            n_blocks = cpu_sample_blocks(cores)
            v, cdv, pv, dt, res, jobs = cpu_port_sample(n_blocks, cores)
            line["cpu_baseline"] = {"value": v, "unit": "candidates/s", "cores": cores, "kind": "port",
                                    "sample": f"{n_blocks} coder-blocks (D=1000, S=36, B=20) of the same workload, C oracle "
                                              f"(oracle/irec_oracle.c), {cores} threads, {dt:.1f} s",
                                    "candidate_dims_per_sec": cdv}
            n_np = max(cores, 16)
            v2, cd2, dt2 = cpu_numpy_port_sample(n_np, cores)
            line["cpu_baselines"] = [line["cpu_baseline"],
                                     {"value": v2, "unit": "candidates/s", "cores": cores, "kind": "port",
                                      "sample": f"{n_np} coder-blocks of the same workload, reference-structured NumPy port "
                                                f"(oracle/ref_numpy.py: [S,B,D] candidate tensor, two log_prob passes, argsort), "
                                                f"{cores} threads, {dt2:.1f} s", "candidate_dims_per_sec": cd2}]
            # index match against the oracle on the same sample blocks (free-running): the sampled blocks are coded again
            line.update(index_match(res, jobs, coder))
            # canonical score vs the reference's own float32 form (oracle/refform_study.py), bounded live sample
            from oracle import refform_study as RS
            rep = RS.run(max(2 * cores, 32), threads=cores)
            line["refform"] = {"blocks": rep["blocks"], "index_match_pct_refform": rep["free_running"]["index_match_pct_refform"],
                               "blocks_identical_pct": rep["free_running"]["blocks_identical_pct"],
                               "teacher_forced_match_pct": rep["teacher_forced"]["teacher_forced_match_pct"],
                               "teacher_forced_partitions": rep["teacher_forced"]["partitions"],
                               "worst_relative_gap_of_a_mismatch": rep["teacher_forced"]["worst_relative_gap_of_a_mismatch"],
                               "n_aux_differs_float32_kl": rep["kl_float32_vs_float64"]["blocks_n_aux_differs_pairwise"],
                               "population": "profiles/r2_refform_study.json (2048 coder-blocks)", "seconds": rep["seconds"]}
            line["index_match_pct_refform"] = rep["free_running"]["index_match_pct_refform"]
            line["teacher_forced_match_pct"] = rep["teacher_forced"]["teacher_forced_match_pct"]
        print(json.dumps(line))
    if world > 1:
        barrier()
        dist.destroy_process_group()


def index_match(res, jobs, coder):
    """GPU == oracle on the CPU sample's coder-blocks: the blocks are coded again in one launch (one block per row)"""
    import torch
    from irec_b200 import engine as E
    dev = torch.device("cuda", torch.cuda.current_device())
    nbk = len(jobs)
    flat = [torch.from_numpy(np.concatenate([j[k] for j in jobs])).to(dev) for k in range(4)]
    offsets, nb, max_dim = E.make_block_offsets(nbk * BLOCK, BLOCK, dev)
    out = E.beam_encode_blocks(*flat, None, offsets, nb, max_dim, OMEGA, S, NBEAMS, SEED)
    smp = out.sample.cpu().numpy()
    match = total = blocks = 0
    for b, r in enumerate(res):
        ref_idx = r["indices"].tolist()
        total += len(ref_idx)
        same = out.indices[b] == ref_idx
        match += len(ref_idx) if same else sum(int(x == y) for x, y in zip(out.indices[b], ref_idx))
        blocks += int(same and np.array_equal(smp[b * BLOCK:(b + 1) * BLOCK].view(np.uint32), r["sample"].view(np.uint32)))
    return {"index_match_pct": 100.0 * match / max(total, 1), "blocks_bit_identical_to_oracle": f"{blocks}/{nbk}"}


if __name__ == "__main__":
    main()
