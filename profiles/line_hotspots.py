#!/usr/bin/env python
"""Maps the per-instruction warp samples of an ncu report to CUDA source lines (via nvdisasm -g line info of the
cubin inside libirec.so) and prints the hottest lines.
usage: python profiles/line_hotspots.py gpurun_out/x.ncu-rep '_Z23k_beam_encode_resident2ILi20E' [cubin-name-substring]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, fn = sys.argv[1], sys.argv[2]
    cub_sub = sys.argv[3] if len(sys.argv) > 3 else "irec_beam."
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "relative-entropy-coding_b200", "lib", "libirec.so")], cwd=tmp,
                   capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.startswith(cub_sub)][0]
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    cur, on, off2line = None, False, {}
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            on = fn in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m:
            off2line[int(m.group(1), 16)] = cur
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > ix["stall_wait"]]
    base = int(data[0][ix["Address"]], 16)
    agg, tot = collections.Counter(), 0.0
    for r in data:
        s = float(r[ix["# Samples"]] or 0)
        tot += s
        agg[off2line.get(int(r[ix["Address"]], 16) - base)] += s
    byfile = collections.Counter()
    for k, v in agg.items():
        byfile[k[0] if k else None] += v
    if os.environ.get("IREC_HOTSPOT_DUMP"):
        import pickle
        pickle.dump((dict(agg), tot), open(os.environ["IREC_HOTSPOT_DUMP"], "wb"))
    print("| file:line | % of warp samples |\n|---|---|")
    for k, v in agg.most_common(40):
        print(f"| {k[0]}:{k[1]} | {100 * v / tot:.2f} |" if k else f"| ? | {100 * v / tot:.2f} |")


if __name__ == "__main__":
    main()
