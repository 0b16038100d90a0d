#!/usr/bin/env python
"""Turns an `ncu --set full` report (gpurun_out/*.ncu-rep) into the markdown summary committed under profiles/.

usage: python profiles/summarize_ncu.py gpurun_out/r2_resident2.ncu-rep "title" > profiles/<name>.md
Needs the `ncu` CLI (no GPU).  Numbers are per launch (first captured launch)."""
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "smsp__inst_executed_op_shared_ld.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, check=True).stdout


def main():
    rep, title = sys.argv[1], sys.argv[2]
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, first = raw[0], raw[1], raw[2]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    print(f"source: `{rep}` (ncu --set full --clock-control none --import-source on), kernel `{first[col['Kernel Name']]}`, "
          f"grid {first[col['Grid Size']]} x block {first[col['Block Size']]}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in RAW_KEYS:
        if k in col:
            print(f"| `{k}` | {first[col[k]]} | {units[col[k]]} |")
    # per-opcode and stall aggregation from the source page
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--launch-count", "1"]))))
    h = src[1]
    ix = {n: i for i, n in enumerate(h)}
    data = [r for r in src[2:] if len(r) > ix["stall_wait"]]

    def f(x):
        try:
            return float(x)
        except ValueError:
            return 0.0

    tot = sum(f(r[ix["# Samples"]]) for r in data) or 1.0
    ops = {}
    for r in data:
        toks = r[ix["Source"]].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")
        a = ops.setdefault(op, [0.0, 0.0, 0.0, 0.0])
        a[0] += f(r[ix["Instructions Executed"]]); a[1] += f(r[ix["L1 Wavefronts Shared"]])
        a[2] += f(r[ix["L1 Wavefronts Shared Ideal"]]); a[3] += f(r[ix["# Samples"]])
    print("\n## warp-instructions by opcode (top 16)\n\n| opcode | executed | shared wavefronts | ideal | % of samples |\n|---|---|---|---|---|")
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:16]:
        print(f"| {k} | {v[0]:.4g} | {v[1]:.4g} | {v[2]:.4g} | {100 * v[3] / tot:.1f} |")
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = {s: sum(f(r[ix[s]]) for r in data) for s in stalls}
    print("\n## warp stall reasons (% of all samples)\n\n| reason | % |\n|---|---|")
    for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]:
        print(f"| {s} | {100 * v / tot:.1f} |")


if __name__ == "__main__":
    main()
