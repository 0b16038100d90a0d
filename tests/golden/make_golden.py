"""Generates tests/golden/beam_golden.json from the CPU oracle (the reference's TensorFlow code cannot be
imported in this environment).  Run: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402
import synth  # noqa: E402

CASES = [
    dict(recipe="c1", D=64, data_seed=0, omega=3., S=36, B=1, seed=42),
    dict(recipe="c1", D=64, data_seed=0, omega=3., S=36, B=20, seed=42),
    dict(recipe="c2", D=1000, data_seed=1, omega=3., S=36, B=20, seed=42),
    dict(recipe="c2", D=192, data_seed=2, omega=3., S=36, B=20, seed=42),
    dict(recipe="c3", D=288, data_seed=3, omega=3., S=20, B=10, seed=7),
]

out = {"note": "oracle outputs (parity unpinned vs TensorFlow); regenerate with make_golden.py", "cases": []}
for c in CASES:
    tl, ts, pl, ps = getattr(synth, c["recipe"])(c["D"], data_seed=c["data_seed"])
    r = O.beam_encode_block(tl, ts, pl, ps, c["omega"], c["S"], c["B"], c["seed"])
    c = dict(c)
    c["indices"] = r["indices"].tolist()
    c["sample_bits"] = r["sample"].view(np.uint32).tolist()
    c["kl"] = r["kl"]
    out["cases"].append(c)
with open(os.path.join(HERE, "beam_golden.json"), "w") as f:
    json.dump(out, f)
print("wrote", len(out["cases"]), "cases")
