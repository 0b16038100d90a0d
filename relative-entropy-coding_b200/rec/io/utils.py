"""write_compressed_code / read_compressed_code -- mirror of rec/io/utils.py:7-216 over the host C++ of libirec.so
(include/irec_io.h: irec_rec_write_file / irec_rec_read_header / irec_rec_unpack).  Files are byte-identical to the
reference's; same signatures, same return values."""
import ctypes as C

import numpy as np

from irec_b200 import native as N


class _Header(C.Structure):
    _fields_ = [("seed", C.c_uint32), ("block_size", C.c_uint32), ("max_index", C.c_uint32), ("image_h", C.c_uint32),
                ("image_w", C.c_uint32), ("image_c", C.c_uint16), ("uses_num_aux_counts_file", C.c_uint16),
                ("uses_index_counts_file", C.c_uint16), ("n_res_blocks", C.c_int32)]


MAX_INDEX_LIMIT = 1 << 28      # alphabet of the arithmetic coder (the reference's examples use max_index = 20)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def write_compressed_code(file_path,
                          seed,
                          image_shape,
                          block_size,
                          block_indices,
                          max_index,
                          num_aux_var_counts_file=None,
                          index_counts_file=None):
    """block_indices: per latent tensor ("residual block") a list of per-coder-block index lists (what
    `coder.encode(...)` returns with block_size set).  rec/io/utils.py:7-106."""
    if len(image_shape) != 3:
        raise ValueError(f"Image shape must be rank 3, but was {image_shape}!")
    if num_aux_var_counts_file is not None:
        # the reference cannot write such a file either: it packs num_aux_var_maxes = [-1] * n with format 'I'
        # (rec/io/utils.py:52,95), which raises struct.error
        raise NotImplementedError("empirical num_aux_var counts are not supported (the reference fails on them too)")
    img_h, img_w, img_c = (int(v) for v in image_shape)
    num_blocks = np.array([len(blk) for blk in block_indices], dtype=np.int32)
    num_aux = np.array([len(b) for blk in block_indices for b in blk], dtype=np.int32)
    flat = [np.asarray(b, dtype=np.int64).reshape(-1) for blk in block_indices for b in blk]
    indices = np.ascontiguousarray(np.concatenate(flat) if flat else np.zeros(0, np.int64), dtype=np.int64)
    index_counts = None
    if index_counts_file is not None:
        index_counts = np.ascontiguousarray(np.load(index_counts_file), dtype=np.int64)
        if index_counts.size != max_index + 1:
            raise ValueError("index counts must have max_index + 1 entries")
    if len(block_indices) > 0xFFFF:
        # the container stores the number of residual blocks as uint16; the reference's struct.pack raises here too
        raise ValueError(f"too many residual blocks for the .rec container: {len(block_indices)} > 65535")
    h = _Header(int(seed), int(block_size), int(max_index), img_h, img_w, img_c, 0, 0, len(block_indices))
    n = C.c_int64(0)
    N.check(N.load_library().irec_rec_write_file(str(file_path).encode(), C.byref(h), _ptr(num_blocks), _ptr(num_aux),
                                                 _ptr(indices), _ptr(index_counts), C.byref(n)), "irec_rec_write_file")
    return int(n.value)


def read_compressed_code(file_path,
                         static_header_size=28,
                         num_aux_var_counts_file=None,
                         index_counts_file=None):
    """-> (seed, image_shape, block_size, block_indices).  rec/io/utils.py:109-216."""
    lib = N.load_library()
    with open(file_path, "rb") as f:
        data = np.frombuffer(f.read(), dtype=np.uint8)
    h = _Header()
    N.check(lib.irec_rec_read_header(_ptr(data), int(data.size), C.byref(h)), "irec_rec_read_header")
    if h.uses_index_counts_file and index_counts_file is None:
        raise ValueError("The compressed file is using empirical index counts, but no counts file was supplied!")
    if h.max_index >= MAX_INDEX_LIMIT:
        raise ValueError(f"corrupt .rec header: max_index {h.max_index} (limit {MAX_INDEX_LIMIT})")
    index_counts = None
    if h.uses_index_counts_file:
        index_counts = np.ascontiguousarray(np.load(index_counts_file), dtype=np.int64)
        if index_counts.size != h.max_index + 1:     # the decoder reads max_index + 1 counts
            raise ValueError(f"index counts file has {index_counts.size} entries, the header needs {h.max_index + 1}")
    num_blocks = np.zeros(max(1, h.n_res_blocks), dtype=np.int32)
    cap_a, cap_i = 1024, 16384
    while True:
        num_aux = np.empty(cap_a, dtype=np.int32)
        indices = np.empty(cap_i, dtype=np.int64)
        na, ni = C.c_int64(0), C.c_int64(0)
        rc = lib.irec_rec_unpack(_ptr(data), int(data.size), _ptr(index_counts), _ptr(num_blocks), _ptr(num_aux), cap_a,
                                 C.byref(na), _ptr(indices), cap_i, C.byref(ni))
        if rc == 0:
            break
        if na.value > cap_a or ni.value > cap_i:
            cap_a, cap_i = max(cap_a, int(na.value)), max(cap_i, int(ni.value))
            continue
        N.check(rc, "irec_rec_unpack")
    block_indices, a0, i0 = [], 0, 0
    for r in range(h.n_res_blocks):
        blocks = []
        for b in range(int(num_blocks[r])):
            k = int(num_aux[a0 + b])
            blocks.append(indices[i0:i0 + k].tolist())
            i0 += k
        a0 += int(num_blocks[r])
        block_indices.append(blocks)
    return int(h.seed), (int(h.image_h), int(h.image_w), int(h.image_c)), int(h.block_size), block_indices
