"""ctypes binding of libirec.so (the C ABI declared in include/irec.h).

PyTorch is used only as the owner of device memory and streams: tensors are handed to the library
as raw device pointers (`tensor.data_ptr()`), together with the current CUDA stream.  There is no
CPU fallback: if the shared library is missing or no CUDA device is present, calls raise.
"""
import collections
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IREC_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libirec.so")

IREC_OK = 0
BLK_OK, BLK_BAD_KL, BLK_TOO_LONG = 0, 1, 2
RECORD_BYTES = 16


class NativeError(RuntimeError):
    """libirec.so is missing / failed; never silently replaced by a CPU path."""


_lib = None

_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

_SIGNATURES = {
    "irec_version": (C.c_int, []),
    "irec_build_hash": (C.c_char_p, []),
    "irec_last_error_string": (C.c_char_p, []),
    "irec_init": (C.c_int, []),
    "irec_get_ndtri_table": (C.c_int, [_vp]),
    "irec_aux_ratio": (C.c_float, [_i32]),
    "irec_set_thread_aux_ratios": (C.c_int, [_vp, _i32]),
    "irec_aux_ratio_len": (C.c_int, []),
    "irec_set_thread_reserved_sms": (C.c_int, [_i32]),
    "irec_tf_op_seed": (C.c_int64, [_i64]),
    "irec_split_permutation": (C.c_int, [_i64, _i64, _vp]),
    "irec_beam_uniform_int": (C.c_int, [_i64, _i64, _i64, _vp, _vp]),
    "irec_is_normal_stream": (C.c_int, [_i64, _i64, _i64, _vp, _vp]),
    "irec_normal_stream_seeded": (C.c_int, [_i64, _i64, _i64, _i64, _vp, _vp]),
    "irec_bm_components_host": (C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    "irec_uniform_int_stream": (C.c_int, [_i64, _i64, _i32, _i32, _i64, _i64, _vp, _vp]),
    "irec_kl_naux": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _vp, _vp, _vp]),
    "irec_schedule": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "irec_beam_encode_workspace_bytes": (C.c_size_t, [_i32, _i64, _i32, _i32, _i32]),
    "irec_beam_encode": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _f32, _i32, _i32, _i64,
                                   _vp, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "irec_beam_encode_path": (C.c_int, [_i32, _i64, _i32, _i32]),
    "irec_beam_decode": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i64, _vp, _i32, _vp, _vp, _vp, _vp]),
    "irec_beam_state_bytes": (C.c_size_t, [_i32, _i32, _i32]),
    "irec_beam_state_init": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _i32, _i32, _i32, _i64, _vp]),
    "irec_beam_state_query": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "irec_beam_step_workspace_bytes": (C.c_size_t, [_i32, _i32]),
    "irec_beam_step_score": (C.c_int, [_vp, _i32, _i32, _i32, _i64, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "irec_beam_step_commit": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _sz, _vp]),
    "irec_beam_state_finish": (C.c_int, [_vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "irec_topb_merge": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "irec_beam_fused_fits": (C.c_int, [_i32, _i32]),
    "irec_beam_fused_workspace_bytes": (C.c_size_t, [_i32, _i32]),
    "irec_beam_encode_fused": (C.c_int, [_vp, _i32, _i32, _i64, _i64, _vp, _i32, _i32, _vp, _sz, _vp]),
    "irec_p2p_exchange_bytes": (C.c_size_t, [_i32, _i32]),
    "irec_p2p_exchange": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "irec_is_workspace_bytes": (C.c_size_t, [_i32]),
    "irec_is_coded_sample": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "irec_is_decode_sample": (C.c_int, [_vp, _vp, _i32, _vp, _i64, _vp, _vp]),
    "irec_is_block_workspace_bytes": (C.c_size_t, [_i32]),
    "irec_is_encode_workspace_bytes": (C.c_size_t, [_i32, _i64, _i64, _i32]),
    "irec_is_encode": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _f32, _i64, _i64,
                                 _vp, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "irec_is_decode": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "irec_launch_count": (C.c_int64, []),
    # include/irec_io.h -- host C++ (arithmetic coder + .rec container), no CUDA calls
    "irec_ac_encode": (C.c_int, [_vp, _i32, _i32, _vp, _i64, _vp, _i64, _vp]),
    "irec_ac_decode": (C.c_int, [_vp, _i32, _i32, _vp, _i64, _vp, _i64, _vp]),
    "irec_rec_pack": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "irec_rec_read_header": (C.c_int, [_vp, _i64, _vp]),
    "irec_rec_unpack": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "irec_rec_write_file": (C.c_int, [C.c_char_p, _vp, _vp, _vp, _vp, _vp, _vp]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))


def load_library():
    """dlopen libirec.so and declare every prototype (no CUDA call is made here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def lib():
    l = load_library()
    if not torch.cuda.is_available():
        raise NativeError("no CUDA device: the iREC coders run on B200 (sm_100a) only; there is no CPU fallback")
    return l


def check(rc, what=""):
    if rc != IREC_OK:
        msg = load_library().irec_last_error_string().decode(errors="replace")
        raise NativeError(f"{what or 'libirec'} failed (code {rc}): {msg}")


def ptr(t):
    """raw device pointer of a CUDA tensor (None -> NULL)"""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("expected a CUDA tensor")
    if not t.is_contiguous():
        raise NativeError("expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(load_library().irec_launch_count())


def tf_op_seed(seed):
    return int(load_library().irec_tf_op_seed(int(seed)))


def aux_ratio(i):
    return float(load_library().irec_aux_ratio(int(i)))


_borrowed = collections.deque()      # (event, tensor): ratio tables still referenced by enqueued kernels


class thread_aux_ratios:
    """`with thread_aux_ratios(table):` -- learned auxiliary variance ratios (a float32 CUDA tensor, or None for the
    power law) for every libirec call made by this thread inside the block (include/irec.h: irec_set_thread_aux_ratios)."""

    def __init__(self, table):
        self.table = table

    def __enter__(self):
        if self.table is not None:
            if self.table.dtype != torch.float32 or self.table.dim() != 1 or self.table.numel() < 1:
                raise NativeError("auxiliary ratio table must be a non-empty 1-D float32 CUDA tensor")
            check(lib().irec_set_thread_aux_ratios(ptr(self.table), int(self.table.numel())), "irec_set_thread_aux_ratios")
        return self

    def __exit__(self, *exc):
        if self.table is not None:
            # the table is borrowed until the enqueued work has finished: instead of synchronising the stream, keep a
            # reference parked behind an event and drop it once the event has completed
            ev = torch.cuda.Event()
            ev.record()
            _borrowed.append((ev, self.table))
            while _borrowed and _borrowed[0][0].query():
                _borrowed.popleft()
            check(load_library().irec_set_thread_aux_ratios(None, 0), "irec_set_thread_aux_ratios")
        return False


class reserved_sms:
    """`with reserved_sms(k):` -- the persistent batch kernels launched by this thread inside the block leave k SMs free, so
    that small kernels of other streams can start while a launch is running (include/irec.h: irec_set_thread_reserved_sms)."""

    def __init__(self, k):
        self.k = int(k)

    def __enter__(self):
        check(load_library().irec_set_thread_reserved_sms(self.k), "irec_set_thread_reserved_sms")
        return self

    def __exit__(self, *exc):
        check(load_library().irec_set_thread_reserved_sms(-1), "irec_set_thread_reserved_sms")
        return False


def split_permutation(n, seed):
    """Coder.split permutation (rec/coding/coder.py:60-67) as a CPU int64 tensor (host C++ in libirec.so)."""
    out = torch.empty(int(n), dtype=torch.int64)
    check(load_library().irec_split_permutation(int(n), int(seed), C.c_void_p(out.data_ptr())), "irec_split_permutation")
    return out


def ndtri_table():
    out = torch.empty(10007, dtype=torch.float32)
    check(lib().irec_get_ndtri_table(C.c_void_p(out.data_ptr())), "irec_get_ndtri_table")
    return out
