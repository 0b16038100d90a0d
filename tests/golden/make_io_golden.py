"""Writes tests/golden/io_golden.json with the REAL reference rec/io (compiled into oracle/_ref by oracle/build_ref.py
from /root/reference; run in this container only):  python tests/golden/make_io_golden.py

Cases: arithmetic-coder code strings for seeded random masses/messages (the shape of rec/io/tests/coding_test.py:9-20
at smaller sizes, plus edge cases), and whole `.rec` files (hex) for index lists shaped like the lossless (S=36 > max_index
is the reference's known quirk, so those use max_index=36) and lossy (max_index=20) examples."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_io  # noqa: E402


def main():
    rng = np.random.Generator(np.random.PCG64(20260917))
    ac_cases = []
    for n_sym, n_msg in [(1, 0), (2, 1), (4, 10), (16, 100), (64, 2000), (21, 500), (37, 300)]:
        P = [1] + [int(v) for v in rng.integers(1, 101, n_sym)]
        msg = [int(v) for v in rng.integers(1, n_sym + 1, n_msg)] + [0]
        ac_cases.append({"P": P, "message": msg})
    ac_cases.append({"P": [1] + [1001] * 20, "message": [20] * 50 + [0]})           # runs of one symbol
    ac_cases.append({"P": [1] + [1001] * 20, "message": [1] * 50 + [0]})
    ac_cases.append({"P": [1, 101, 101, 101], "message": [3, 2, 1, 3, 0]})
    ac_cases.append({"P": [5, 1, 1, 1, 90], "message": [4] * 200 + [1, 2, 3, 0]})    # skewed masses
    codes = ref_io.call([{"op": "ac_encode", **c} for c in ac_cases])
    decoded = ref_io.call([{"op": "ac_decode_fast", "P": c["P"], "code": code} for c, code in zip(ac_cases, codes)])
    for c, code, dec in zip(ac_cases, codes, decoded):
        assert dec == c["message"]
        c["code"] = code
    rec_cases = []
    tmp = tempfile.mkdtemp()
    shapes = [((32, 32, 3), 1000, 36, [(9, 36, 12)] * 3), ((512, 768, 3), 1000, 20, [(13, 20, 40), (30, 20, 25)]),
              ((8, 8, 1), 0, 5, [(1, 5, 1)]), ((16, 16, 3), 50, 20, [(4, 20, 0), (2, 20, 3)])]
    for k, (shape, bs, max_index, tensors) in enumerate(shapes):
        bi = []
        for nblk, S, mean_aux in tensors:
            blocks = []
            for _ in range(nblk):
                n_aux = int(rng.integers(max(0, mean_aux - 3), mean_aux + 4)) if mean_aux else int(rng.integers(0, 2))
                blocks.append([int(v) for v in rng.integers(0, S, n_aux)])
            bi.append(blocks)
        path = os.path.join(tmp, f"g{k}.rec")
        req = {"op": "rec_write", "path": path, "seed": 42 + k, "image_shape": list(shape), "block_size": bs,
               "block_indices": bi, "max_index": max_index}
        size, back = ref_io.call([req, {"op": "rec_read", "path": path}])
        assert back["block_indices"] == bi and back["seed"] == 42 + k
        rec_cases.append({"seed": 42 + k, "image_shape": list(shape), "block_size": bs, "max_index": max_index,
                          "block_indices": bi, "file_hex": open(path, "rb").read().hex()})
        assert size == len(rec_cases[-1]["file_hex"]) // 2
    out = {"note": "written by the reference's own rec/io (oracle/_ref); regenerate with tests/golden/make_io_golden.py",
           "ac": ac_cases, "rec": rec_cases}
    with open(os.path.join(ROOT, "tests", "golden", "io_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(ac_cases), "coder cases and", len(rec_cases), "files")


if __name__ == "__main__":
    main()
