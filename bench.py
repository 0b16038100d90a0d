#!/usr/bin/env python
"""bench.py -- throughput of the iREC beam-search encode hot path on B200.

Workload (BASELINE.json configs[3] shape, weak scaling): per GPU a batch of IMAGES_PER_GPU ImageNet32-shaped
synthetic images through resnet_vae latents -- 24 latent tensors [16,16,32] per image, block_size 1000
(9 coder-blocks per tensor), beam search n_beams=20, extra_samples=1.2 (S=36), kl_per_partition=3 nats,
coding seed 42.  8 GPUs x 128 images = the 1024-image batch of configs[3].  One "step" = one pass of the
encoder over the whole per-GPU batch: 24 sequential level launches (a level's prior depends on the previous
level's sample in the real model), every launch coding 128 x 9 blocks.

  value  : candidates scored / s, inputs resident in HBM (CUDA events)
  e2e    : same through the public API (rec.coding.BeamSearchCoder.encode_batch) from pinned host buffers,
           host<->device copies and index read-back inside the timed region
  roofline: INT32/FP32 issue rate of the dominant kernel (k_beam_encode_resident2<20>), algorithmic
           instructions per candidate-dim W = 10 + 24/B' (SURVEY.md 8d); secondary: HBM bytes and the
           shared-memory wavefront rate (the measured limiter of the hot loop, profiles/)

`--impl reference` times the CPU port of the reference (the C oracle, all host threads) on a bounded sample.
"""
import argparse
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

IMAGES_PER_GPU = 128
LEVELS = 24
LATENT = 8192            # 16*16*32
BLOCK = 1000
OMEGA, EXTRA, NBEAMS, SEED = 3.0, 1.2, 20, 42
S = int(np.exp(OMEGA * EXTRA))
MAX_AUX = 256
METRIC = "iREC candidates scored/sec (beam-search encode; candidate = one (sample, beam) pair scored over all dims of its coder-block)"


def synth_level(image, level, n=LATENT):
    """SURVEY.md 8(d) C2 recipe, one latent tensor"""
    import synth
    return synth.c2(n, data_seed=1000 * image + level)


def work_model(n_aux, dims):
    """algorithmic work of a set of blocks: candidates, candidate-dims, partitions, lane-instructions"""
    n_aux = np.asarray(n_aux, np.int64)
    dims = np.asarray(dims, np.int64)
    cand = S * 1 + (n_aux - 1) * S * NBEAMS                 # t = 0 has one (empty) beam
    cd = cand * dims
    instr = dims * (S * 1 * (10 + 24 / 1) + (n_aux - 1) * S * NBEAMS * (10 + 24 / NBEAMS))
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


def cpu_port_sample(n_blocks, threads):
    """oracle (C port of the reference) on `n_blocks` coder-blocks of the same workload, `threads` host threads"""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.lib()
    O.ndtri_table()
    perm = O.shuffle_perm(LATENT, SEED)
    jobs = []
    for i in range(n_blocks):
        tl, ts, pl, ps = synth_level(i // 8, i % LEVELS)
        sel = perm[(i % 8) * BLOCK:(i % 8 + 1) * BLOCK]
        jobs.append((tl[sel], ts[sel], pl[sel], ps[sel]))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(lambda j: O.beam_encode_block(*j, OMEGA, S, NBEAMS, SEED), jobs))
    dt = time.perf_counter() - t0
    cand, cd, parts, _ = work_model([r["n_aux"] for r in res], [BLOCK] * n_blocks)
    return cand / dt, cd / dt, parts / dt, dt, res, jobs


def cpu_sample_blocks(cores):
    """bounded CPU sample: ~0.4 s of single-core work per coder-block -> 10-20 s per step on all cores"""
    return int(min(1024, max(32 * cores, 64)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_blocks = cpu_sample_blocks(cores)
    vals = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
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
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "sample": f"{n_blocks} coder-blocks of D=1000 per step"},
        "candidate_dims_per_sec": float(np.mean([x[1] for x in vals])),
        "partitions_per_sec": float(np.mean([x[2] for x in vals])),
        "cpu_baseline": {"value": v, "unit": "candidates/s", "cores": cores, "kind": "port",
                         "sample": f"{n_blocks} coder-blocks (D=1000, S=36, B=20) per step, C oracle, {cores} threads"},
        "e2e": {"value": v, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "TensorFlow 2.1/TFP 0.9 are not installable here; this is the C port of the reference algorithm "
                "(oracle/irec_oracle.c) run block-parallel on all host threads",
    }
    print(json.dumps(line))


def workload_name():
    return (f"C4-shaped: {IMAGES_PER_GPU} images/GPU x {LEVELS} resnet_vae latents [16,16,32], block_size {BLOCK}, "
            f"beam_search n_beams={NBEAMS} extra_samples={EXTRA} (S={S}) kl_per_partition={OMEGA}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--images-per-gpu", type=int, default=IMAGES_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="kernel-tuning runs: skip the end-to-end leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    from irec_b200 import engine as E, native as N, Normal
    from rec.coding import BeamSearchCoder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_img = args.images_per_gpu

    # ---------------- synthetic inputs (pinned host + device copies) ----------------
    host = []      # per level: 4 pinned tensors [n_img, LATENT]
    for lvl in range(LEVELS):
        arrs = [synth_level(rank * n_img + i, lvl) for i in range(n_img)]
        host.append([torch.from_numpy(np.stack([a[k] for a in arrs])).pin_memory() for k in range(4)])
    devt = [[t.to(dev, non_blocking=True) for t in lv] for lv in host]
    coder = BeamSearchCoder(kl_per_partition=OMEGA, n_beams=NBEAMS, extra_samples=EXTRA, block_size=BLOCK)
    perm = coder._permutation(LATENT, SEED, dev)
    gather = (perm[None, :] + torch.arange(n_img, device=dev)[:, None] * LATENT).reshape(-1).contiguous()
    offsets, nb, max_dim = E.make_block_offsets(LATENT, BLOCK, dev, n_items=n_img)
    dims = (offsets[1:] - offsets[:-1]).cpu().numpy()
    lib = N.lib()
    ws_bytes = int(lib.irec_beam_encode_workspace_bytes(nb, max_dim, S, NBEAMS, MAX_AUX))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    out_idx = [torch.empty((nb, MAX_AUX), dtype=torch.int32, device=dev) for _ in range(LEVELS)]
    out_na = [torch.empty(nb, dtype=torch.int32, device=dev) for _ in range(LEVELS)]
    out_st = [torch.empty(nb, dtype=torch.int32, device=dev) for _ in range(LEVELS)]
    out_sample = [torch.empty(n_img * LATENT, dtype=torch.float32, device=dev) for _ in range(LEVELS)]
    stream = N.stream_ptr()

    def launch_level(lvl):
        tl, ts, pl, ps = (t.reshape(-1) for t in devt[lvl])
        N.check(lib.irec_beam_encode(N.ptr(tl), N.ptr(ts), N.ptr(pl), N.ptr(ps), N.ptr(gather), N.ptr(offsets), nb,
                                     max_dim, OMEGA, S, NBEAMS, SEED, N.ptr(out_idx[lvl]), MAX_AUX, N.ptr(out_na[lvl]),
                                     N.ptr(out_st[lvl]), N.ptr(out_sample[lvl]), N.ptr(ws), ws_bytes, stream),
                "irec_beam_encode")

    def step_resident(events=None):
        for lvl in range(LEVELS):
            if events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            launch_level(lvl)
            if events is not None:
                e1.record()
                events.append((e0, e1))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = N.launch_count()
    kernel_events = []
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        step_resident(kernel_events)
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
    kernel_ms = np.array([a.elapsed_time(b) for a, b in kernel_events]).reshape(args.steps, LEVELS)
    kernel_ms_avg = float(kernel_ms.mean())
    achieved_instr = float(np.mean(per_level_instr)) / (kernel_ms_avg * 1e-3)

    # ---------------- end-to-end through the public API (host buffers) ----------------
    sample_host = [torch.empty((n_img, LATENT), dtype=torch.float32, pin_memory=True) for _ in range(LEVELS)]

    def step_e2e():
        """all levels through the public API from pinned host buffers; the index lists of level l are read back and built
        while level l + 1 runs (encode_batch(lazy=True)); every level's inputs cross H2D and its sample and indices D2H"""
        total, pending = 0, None
        for lvl in range(LEVELS):
            tl, ts, pl, ps = (t.to(dev, non_blocking=True) for t in host[lvl])
            get_indices, sample = coder.encode_batch(Normal(tl, ts), Normal(pl, ps), seed=SEED, lazy=True)
            sample_host[lvl].copy_(sample, non_blocking=True)
            if pending is not None:
                total += sum(len(b) for img in pending() for b in img)
            pending = get_indices
        total += sum(len(b) for img in pending() for b in img)
        torch.cuda.synchronize()
        return total, sample_host[-1]

    e2e_s = float("nan")
    if not args.no_e2e:
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 2))
        for _ in range(e2e_steps):
            parts_e2e, _ = step_e2e()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        assert parts_e2e == parts, (parts_e2e, parts)
    h2d = LEVELS * 4 * n_img * LATENT * 4
    d2h = LEVELS * (n_img * LATENT * 4 + nb * (2 + int(na_all.max())) * 4)

    # ---------------- max over ranks, sums over ranks ----------------
    t_res = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([cand, cd, parts], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms = float(t_res[0]), float(t_res[1])
    cand_all, cd_all, parts_all = (float(x) for x in tot)
    sec = ms_total * 1e-3 / args.steps

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        peak_instr = sm_count * 128 * sm_max_mhz * 1e6
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes_level = float(16 * dims.sum() + 4 * dims.sum() + 4 * na_all.mean(axis=0).sum())
        # measured per-launch DRAM traffic and shared-memory wavefronts per candidate-dim of the same launch shape,
        # from the committed `ncu --set full` capture (profiles/r1_bench_launch_ncu.json)
        traffic, traffic_src, wf_per_cd = None, None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_launch_ncu.json")))
            if prof.get("images_per_gpu") == n_img:
                traffic = float(prof["dram_bytes_read"] + prof["dram_bytes_write"])
            wf_per_cd = float(prof["shared_wavefronts"]) / float(prof["candidate_dims"])
            traffic_src = prof.get("source")
        except Exception:
            pass
        smem_rate = (cd_all / world / sec) * wf_per_cd if wf_per_cd else float("nan")
        line = {
            "metric": METRIC, "value": cand_all / sec, "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "images_total": n_img * world, "parallelism": f"dp{world} (images sharded, no collective in the loop)",
                       "l2": "inputs of one step (403 MB/GPU) exceed the 126 MB L2"},
            "candidate_dims_per_sec": cd_all / sec, "partitions_per_sec": parts_all / sec,
            "index_match_pct": None,
            "gpu_launches": int(launches),
            "clocks": clock_info,
            "e2e": {"value": cand_all / (e2e_ms * 1e-3), "unit": "candidates/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms},
            "roofline": {"bound": "issue", "kernel": "k_beam_encode_resident2<20>", "achieved": achieved_instr / 1e9,
                         "peak": peak_instr / 1e9, "unit": "G lane-instr/s", "frac": achieved_instr / peak_instr,
                         "peak_source": f"{sm_count} SMs x 128 lanes x {sm_max_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz)",
                         "work_model": "W = 10 + 24/B' lane-instr per candidate-dim (SURVEY.md 8d)",
                         "avg_launch_ms": kernel_ms_avg, "traffic": traffic, "traffic_source": traffic_src,
                         "smem": {"wavefronts_per_candidate_dim": wf_per_cd, "achieved_gwf_s": smem_rate / 1e9,
                                  "peak_gwf_s": sm_count * sm_max_mhz * 1e6 / 1e9,
                                  "frac": smem_rate / (sm_count * sm_max_mhz * 1e6), "source": traffic_src},
                         "hbm": {"achieved_gbs": alg_bytes_level / (kernel_ms_avg * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "frac": alg_bytes_level / (kernel_ms_avg * 1e-3) / 1e9 / hbm_peak,
                                 "algorithmic_bytes_per_launch": alg_bytes_level}},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_blocks = cpu_sample_blocks(cores)
            v, cdv, pv, dt, res, jobs = cpu_port_sample(n_blocks, cores)
            line["cpu_baseline"] = {"value": v, "unit": "candidates/s", "cores": cores, "kind": "port",
                                    "sample": f"{n_blocks} coder-blocks (D=1000, S=36, B=20) of the same workload, C oracle "
                                              f"(oracle/irec_oracle.c), {cores} threads, {dt:.1f} s",
                                    "candidate_dims_per_sec": cdv}
            # index match against the oracle on the same sample blocks (free-running)
            match, total = 0, 0
            idx0 = [o.cpu().numpy() for o in out_idx]
            na0 = [o.cpu().numpy() for o in out_na]
            for i, r in enumerate(res):
                img, lvl, b = i // 8, i % LEVELS, i % 8
                if img >= n_img:
                    continue
                blk = img * 9 + b
                got = idx0[lvl][blk, :na0[lvl][blk]]
                total += len(r["indices"])
                match += int((got == r["indices"]).sum()) if len(got) == len(r["indices"]) else 0
            line["index_match_pct"] = 100.0 * match / max(total, 1)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
