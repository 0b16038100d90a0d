"""Code-generation canary for the hot loop of k_beam_encode_resident2<20> (no GPU needed: reads the SASS of the built
libirec.so).  ptxas sits close to the 168-register cap of a 384-thread CTA there, and twice during development an
unrelated edit elsewhere in the kernel made it serialise the quantile gathers (LDS followed directly by its use:
39.0 ms per launch instead of 35.7 ms).  This test fails if that happens again:
the median distance (in instructions) between a table gather and the first use of its result must stay >= 10, and the
loop body must not touch local memory more than a couple of times."""
import os
import re
import shutil
import statistics
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "relative-entropy-coding_b200", "lib", "libirec.so")


def _kernel_sass(tmp):
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True, check=True)
    cub = [f for f in os.listdir(tmp) if f.startswith("irec_beam.")][0]
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True, check=True).stdout
    out, on, cur = [], False, None
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            on = "k_beam_encode_resident2ILi20E" in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = int(m.group(2)) if m.group(1).endswith("irec_resident2.cuh") else -1
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m:
            out.append((cur, m.group(1).strip()))
    return out


def test_quantile_gathers_stay_software_pipelined(built):
    if not (shutil.which("cuobjdump") and shutil.which("nvdisasm")):
        pytest.skip("CUDA binary utilities not available")
    src = open(os.path.join(ROOT, "relative-entropy-coding_b200", "csrc", "irec_resident2.cuh")).read().splitlines()
    gather_line = 1 + next(i for i, l in enumerate(src) if "tv[g][e] = *reinterpret_cast<const float*>(T2b + (ad[k][e] + cb[g]));" in l)
    with tempfile.TemporaryDirectory() as tmp:
        ins = _kernel_sass(tmp)
    # unrolled loop bodies = long runs of instructions attributed to the scoring chunk (source lines of r2_score_chunk);
    # the kernel holds several instances (table / in-place exponents, 1 / 10 beams per pass): the hot one is the
    # table-driven pass of 10 beams x 3 samples = the run with >= 100 gathers whose neighbourhood loads 8-byte exponent-table entries (and has no MATCH of the in-place bank assignment)
    lo, hi = gather_line - 60, gather_line + 45
    runs, i = [], 0
    while i < len(ins):
        if lo <= ins[i][0] <= hi:
            j = i
            while j < len(ins) and (lo <= ins[j][0] <= hi or ins[j][0] in range(28, 60)):
                j += 1
            if j - i > 200:
                runs.append((ins[i:j], ins[max(0, i - 120):i] + ins[j:j + 120]))
            i = j
        else:
            i += 1
    hot = []
    for body, around in runs:
        dist = []
        for k, (line, text) in enumerate(body):
            m = re.match(r"LDS R(\d+), ", text)
            if not (m and line == gather_line):
                continue
            reg = re.compile(r"\bR" + m.group(1) + r"\b")
            for n in range(k + 1, len(body)):
                ops = body[n][1].split(",", 1)
                if len(ops) > 1 and reg.search(ops[1]):
                    dist.append(n - k)
                    break
        if len(dist) >= 100 and any("LDG.E.64" in t for _, t in body + around) and not any("MATCH" in t for _, t in body + around):      # 8-byte exponent-table entries
            hot.append((dist, sum(1 for _, t in body if re.match(r"(LDL|STL)", t))))
    assert len(hot) == 1, [len(d) for d, _ in hot]
    dist, local = hot[0]
    assert statistics.median(dist) >= 10, (statistics.median(dist), sorted(dist)[:10])
    assert local <= 4, local


def _tmem_kernel_sass(tmp):
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True, check=True)
    cub = [f for f in os.listdir(tmp) if f.startswith("irec_tmem.")][0]
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True, check=True).stdout
    out, on, cur = [], False, None
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            on = "k_beam_encode_tmemILi20E" in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = int(m.group(2)) if m.group(1).endswith("irec_tmem.cu") else -1
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m:
            out.append((cur, m.group(1).strip()))
    return out


def test_tmem_kernel_hot_loop(built):
    """k_beam_encode_tmem<20> (the default batch kernel): the scoring loop reads its beams / coefficients with tensor-memory
    loads (LDTM), keeps the quantile gathers software-pipelined (median gather -> use distance >= 10 instructions) and does
    not touch local memory; the kernel allocates tensor memory (UTCATOMSWS / tcgen05.alloc) and stores to it (STTM)."""
    if not (shutil.which("cuobjdump") and shutil.which("nvdisasm")):
        pytest.skip("CUDA binary utilities not available")
    src = open(os.path.join(ROOT, "relative-entropy-coding_b200", "csrc", "irec_tmem.cu")).read().splitlines()
    gather_line = 1 + next(i for i, l in enumerate(src) if "tv[g][e] = *reinterpret_cast<const float*>(T2b + (ad[k][e] + cb[b0 + g]));" in l)
    with tempfile.TemporaryDirectory() as tmp:
        ins = _tmem_kernel_sass(tmp)
    assert any(t.startswith("LDTM") for _, t in ins) and any(t.startswith("STTM") for _, t in ins)
    idx = [i for i, (c, t) in enumerate(ins) if c == gather_line and re.match(r"LDS R\d+", t)]
    assert idx
    groups, g = [], [idx[0]]
    for a, b in zip(idx, idx[1:]):
        if b - a > 400:
            groups.append(g)
            g = []
        g.append(b)
    groups.append(g)
    hot = max(groups, key=len)                       # the main scoring rounds: 4 sample groups x 5 beams x 4 dims = 80 gathers per quad
    assert len(hot) >= 60, [len(x) for x in groups]
    body = ins[hot[0]:hot[-1] + 1]
    assert not any(re.match(r"(LDL|STL)", t) for _, t in body)
    around = ins[max(0, hot[0] - 120):hot[-1] + 120]
    assert sum(1 for _, t in around if t.startswith("LDTM")) >= 8          # 3 coefficient quads + 5 beam quads per quad of dims
    dist = []
    for k in hot:
        m = re.match(r"LDS R(\d+),", ins[k][1])
        reg = re.compile(r"\bR" + m.group(1) + r"\b")
        for n in range(k + 1, min(k + 600, len(ins))):
            ops = ins[n][1].split(",", 1)
            if len(ops) > 1 and reg.search(ops[1]):
                dist.append(n - k)
                break
    assert statistics.median(dist) >= 10, statistics.median(dist)
