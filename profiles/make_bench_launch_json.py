#!/usr/bin/env python
"""profiles/<round>_bench_launch_ncu.json: per-launch DRAM traffic and shared-memory wavefronts of the dominant kernel at the
default bench.py launch shape, from one `ncu --set full` capture, joined with the work of the captured level.

  IREC_BENCH_DUMP=gpurun_out/bench_levels.json ncu --set full --clock-control none --import-source on \\
      -k regex:k_beam_encode_resident2 -s <skip> -c 1 -o gpurun_out/<name> python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline
  python profiles/make_bench_launch_json.py gpurun_out/<name>.ncu-rep gpurun_out/bench_levels.json <skip> [out.json]

<skip> launches of the kernel precede the captured one; bench.py launches it once per level (24 levels per step)."""
import csv
import io
import json
import os
import subprocess
import sys


def main():
    rep, dump, skip = sys.argv[1], sys.argv[2], int(sys.argv[3])
    out_name = sys.argv[4] if len(sys.argv) > 4 else "r1_bench_launch_ncu.json"
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                                                     check=True).stdout)))
    hdr, units, row = raw[0], raw[1], raw[2]
    col = {h: i for i, h in enumerate(hdr)}

    def val(k, scale_units=True):
        v = float(row[col[k]].replace(",", ""))
        u = units[col[k]].lower()
        if scale_units:
            for pre, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3)):
                if u.startswith(pre):
                    v *= m
        return v

    levels = json.load(open(dump))
    lvl = levels["levels"][skip % len(levels["levels"])]
    out = {
        "source": f"profiles/{os.path.basename(rep).replace('.ncu-rep', '')}: ncu --set full --clock-control none, kernel {row[col['Kernel Name']]}, "
                  f"launch #{skip} of `python bench.py --steps 1 --warmup 1` (level {skip % len(levels['levels'])})",
        "kernel": row[col["Kernel Name"]].replace("void ", "").split("(")[0].replace("(int)", ""),
        "images_per_gpu": levels["images_per_gpu"],
        "gpu_time_ms": val("gpu__time_duration.sum") / (1e6 if units[col["gpu__time_duration.sum"]].startswith("ns") else
                                                         1e3 if units[col["gpu__time_duration.sum"]].startswith("us") else 1.0),
        "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
        "shared_wavefronts": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "candidate_dims": lvl["candidate_dims"], "candidates": lvl["candidates"], "partitions": lvl["partitions"],
    }
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), out_name), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
