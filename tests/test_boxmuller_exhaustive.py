"""Exhaustive check of the table-driven float64 log / sincos behind the Box-Muller candidates
(csrc/irec_boxmuller.cuh) against the definition the oracle uses ("float64 libm, round once",
oracle/irec_oracle.c box_muller): the arguments are 23-bit integers, so ALL 2^23 values of each of the three functions
are compared, bit for bit.  The product side is the HOST build of the very functions the kernels run (same IEEE
fma / mul / add sequence; irec_bm_components_host)."""
import ctypes as C

import numpy as np

from irec_b200 import native as N
from oracle import oracle as O


def test_all_2_23_arguments_bit_identical(built):
    lib = N.load_library()
    olib = O.lib()
    olib.orc_bm_components.restype = None
    olib.orc_bm_components.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    n = 1 << 23
    chunk = 1 << 21
    bad = {"logf": [], "sin": [], "cos": []}
    for first in range(0, n, chunk):
        m = np.arange(first, first + chunk, dtype=np.uint32)
        got = [np.empty(chunk, np.float32) for _ in range(3)]
        ref = [np.empty(chunk, np.float32) for _ in range(3)]
        N.check(lib.irec_bm_components_host(C.c_void_p(m.ctypes.data), chunk, *[C.c_void_p(a.ctypes.data) for a in got]),
                "irec_bm_components_host")
        olib.orc_bm_components(C.c_void_p(m.ctypes.data), chunk, *[C.c_void_p(a.ctypes.data) for a in ref])
        for name, g, r in zip(("logf", "sin", "cos"), got, ref):
            diff = np.nonzero(g.view(np.uint32) != r.view(np.uint32))[0]
            bad[name].extend((int(first + i), float(g[i]), float(r[i])) for i in diff[:8])
    assert bad == {"logf": [], "sin": [], "cos": []}, bad
