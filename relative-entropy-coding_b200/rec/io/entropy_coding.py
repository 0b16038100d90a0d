"""ArithmeticCoder -- mirror of the reference's Cython class (rec/io/entropy_coding.pyx:19-302) over the host C++ of
libirec.so (include/irec_io.h: irec_ac_encode / irec_ac_decode).  Same constructor, same methods, same results: a code is
a list of '0'/'1' characters, a decoded message a list of ints ending with the end-of-message symbol 0."""
import ctypes as C

import numpy as np

from irec_b200 import native as N


class ArithmeticCoder(object):

    def __init__(self, P, precision=32):
        P = np.asarray(P)
        self._P = P
        self._precision = int(precision)
        self._counts = np.ascontiguousarray(P, dtype=np.int64)
        if self._counts.ndim != 1 or self._counts.size == 0:
            raise ValueError("P must be a non-empty vector of symbol masses")
        # the reference's public attributes (entropy_coding.pyx:27-47): cumulative masses and their total
        self.D = np.cumsum(self._counts)
        self.C = self.D - self._counts
        self.R = int(self.D[-1])

    def _ptr(self, a):
        return C.c_void_p(a.ctypes.data)

    def encode(self, message):
        """entropy_coding.pyx:51-117"""
        lib = N.load_library()
        msg = np.ascontiguousarray(np.asarray(message), dtype=np.int64).reshape(-1)
        cap = 64 * (msg.size + 2) + 256
        while True:
            out = np.empty(cap, dtype=np.uint8)
            n = C.c_int64(0)
            rc = lib.irec_ac_encode(self._ptr(self._counts), int(self._counts.size), self._precision, self._ptr(msg),
                                    int(msg.size), self._ptr(out), cap, C.byref(n))
            if rc == 0:
                break
            if n.value > cap:
                cap = int(n.value)
                continue
            N.check(rc, "irec_ac_encode")
        return [chr(48 + int(b)) for b in out[:n.value]]

    def decode(self, code):
        """entropy_coding.pyx:121-208 (linear symbol scan in the reference; same result as decode_fast)"""
        return self.decode_fast(code)

    def decode_fast(self, code, verbose=False):
        """entropy_coding.pyx:212-302"""
        lib = N.load_library()
        if isinstance(code, str):
            bits = np.frombuffer(code.encode("ascii"), dtype=np.uint8) - 48
        else:
            bits = np.array([1 if c in ("1", 1, True) else 0 for c in code], dtype=np.uint8)
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        cap = max(1024, 2 * bits.size + 16)
        while True:
            out = np.empty(cap, dtype=np.int64)
            n = C.c_int64(0)
            rc = lib.irec_ac_decode(self._ptr(self._counts), int(self._counts.size), self._precision, self._ptr(bits),
                                    int(bits.size), self._ptr(out), cap, C.byref(n))
            if rc == 0:
                break
            if n.value > cap:
                cap = int(n.value)
                continue
            N.check(rc, "irec_ac_decode")
        return [int(v) for v in out[:n.value]]
