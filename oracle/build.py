"""Build recipe for the CPU oracle (test infrastructure only).

Compiles oracle/irec_oracle.c into oracle/_build/libirec_oracle.so with strict IEEE flags
(no fma contraction, no fast-math).  The reference itself (TensorFlow 2.1 + TFP 0.9 Python) cannot
be compiled or imported here, so there is no oracle/_ref; see DESIGN.md.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libirec_oracle.so")
SRC = os.path.join(HERE, "irec_oracle.c")


FLAGS = ["-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fexcess-precision=standard", "-Wall"]


def _want() -> str:
    import hashlib
    with open(SRC, "rb") as f:
        return hashlib.sha256(" ".join(FLAGS).encode() + f.read()).hexdigest()


def build(force: bool = False) -> str:
    """freshness by content (source + flags hash beside the library), not by mtime"""
    os.makedirs(OUT_DIR, exist_ok=True)
    want, stamp = _want(), LIB + ".hash"
    if (not force) and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return LIB
    subprocess.check_call(["gcc"] + FLAGS + ["-o", LIB, SRC, "-lm"])
    with open(stamp, "w") as f:
        f.write(want)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
