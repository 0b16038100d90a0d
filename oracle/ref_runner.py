"""Subprocess side of oracle/ref_io.py: runs the REAL reference rec.io (compiled into oracle/_ref by build_ref.py).

Reads one JSON request on stdin, writes one JSON reply on stdout.  Must run in its own interpreter with oracle/_ref first
on sys.path, because the product package is also called `rec`.  Test infrastructure only."""
import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_ref"))


def main():
    req = json.load(sys.stdin)
    import numpy as np
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):      # the reference prints and shows tqdm bars
        from rec.io.entropy_coding import ArithmeticCoder
        from rec.io import utils as U
        out = []
        for r in req["requests"]:
            op = r["op"]
            if op == "ac_encode":
                ac = ArithmeticCoder(np.array(r["P"], dtype=np.int32), precision=r.get("precision", 32))
                out.append("".join(ac.encode(np.array(r["message"], dtype=np.int64))))
            elif op in ("ac_decode", "ac_decode_fast"):
                ac = ArithmeticCoder(np.array(r["P"], dtype=np.int32), precision=r.get("precision", 32))
                fn = ac.decode if op == "ac_decode" else ac.decode_fast
                out.append([int(v) for v in fn(r["code"])])
            elif op == "rec_write":
                U.write_compressed_code(r["path"], r["seed"], tuple(r["image_shape"]), r["block_size"],
                                        [[np.array(b, dtype=np.int64) for b in blk] for blk in r["block_indices"]], r["max_index"])
                out.append(os.path.getsize(r["path"]))
            elif op == "rec_read":
                seed, shape, bs, bi = U.read_compressed_code(r["path"])
                out.append({"seed": int(seed), "image_shape": [int(v) for v in shape], "block_size": int(bs),
                            "block_indices": [[[int(v) for v in b] for b in blk] for blk in bi]})
            else:
                raise ValueError(op)
    json.dump({"replies": out}, sys.stdout)


if __name__ == "__main__":
    main()
