"""Build recipe for oracle/_ref: the reference's own rec/io (arithmetic coder + .rec container), compiled where it lies.

The reference's coding hot path (rec/coding) is TensorFlow 2.1 Python and cannot be built here, but its IO side
branch -- the only native component of the reference -- can: `rec/io/entropy_coding.pyx` is Cython, `data_structures.py`
and `utils.py` are plain Python that Cython also compiles.  This recipe cythonizes the three files FROM /root/reference
(nothing is copied into the repository; only the compiled extension modules land in oracle/_ref/rec/io/, which is
git-ignored) with the one directive the 2020 source needs under Cython 3 (`cpow=True`: `2**precision` must stay an
integer power, rec/io/entropy_coding.pyx:61-63).  Test infrastructure only: imported through oracle/ref_io.py by tests/
(always in a subprocess, because the product package is also called `rec`).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_IO = "/root/reference/rec/io"
OUT = os.path.join(HERE, "_ref")
PKG = os.path.join(OUT, "rec", "io")
MODULES = [("entropy_coding", "entropy_coding.pyx"), ("data_structures", "data_structures.py"), ("utils", "utils.py")]


def available() -> bool:
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(PKG, m + ext)) for m, _ in MODULES)


def build(force: bool = False) -> bool:
    """returns True when oracle/_ref is usable (already built, or built now from /root/reference)"""
    if available() and not force:
        return True
    if not os.path.isdir(REF_IO):
        return False                      # GPU box: only the prebuilt files travel
    import numpy
    os.makedirs(PKG, exist_ok=True)
    import tempfile
    tmp = tempfile.mkdtemp(prefix="irec_ref_c_")      # generated C stays out of the repository tree
    for d in (os.path.join(OUT, "rec"), PKG):
        init = os.path.join(d, "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()       # empty package markers (the reference's are empty too)
    inc = sysconfig.get_paths()["include"]
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    for mod, src in MODULES:
        c_file = os.path.join(tmp, mod + ".c")
        subprocess.check_call([sys.executable, "-m", "cython", "-3", "-X", "cpow=True", "-o", c_file, os.path.join(REF_IO, src)])
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-w", "-I", inc, "-I", numpy.get_include(), c_file,
                               "-o", os.path.join(PKG, mod + ext)])
    return available()


if __name__ == "__main__":
    print("oracle/_ref available:", build(force="--force" in sys.argv))
