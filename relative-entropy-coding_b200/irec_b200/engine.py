"""Flat-tensor engine over libirec.so: everything the rec.coding classes need, on CUDA tensors.

All functions take float32 CUDA tensors (flattened) plus the block structure of Coder.split
(rec/coding/coder.py:38-85): `gather_idx` (int64 CUDA tensor or None) and `block_offsets` (int64
CUDA tensor [nb+1]).  Results needed on the host (index lists) cost exactly one device->host copy.
"""
import ctypes as C
import os

import torch

from . import native as N
from . import sharding as SH


class CodingError(Exception):
    """Basis exception class for errors occurring in rec.coding (rec/coding/utils.py:4-7)."""


def _f32c(t, device):
    if str(device).startswith("cuda") and not torch.cuda.is_available():
        raise N.NativeError("no CUDA device: the iREC coders run on B200 (sm_100a) only; there is no CPU fallback")
    if not isinstance(t, torch.Tensor):
        if hasattr(t, "__dlpack__") and hasattr(t, "__dlpack_device__"):
            t = torch.from_dlpack(t)               # zero-copy for CUDA producers (tf eager / cupy / jax arrays)
        else:
            if hasattr(t, "numpy"):
                t = t.numpy()
            t = torch.as_tensor(t)
    return t.to(device=device, dtype=torch.float32).contiguous()


def make_block_offsets(n_total, block_size, device, n_items=1):
    """offsets of consecutive `block_size` chunks (last one short), repeated for n_items tensors of
    n_total dims each (coder.py:70-81)."""
    per = list(range(0, n_total, block_size)) if block_size else [0]
    offs = []
    for it in range(n_items):
        offs.extend(it * n_total + o for o in per)
    offs.append(n_items * n_total)
    max_dim = min(block_size, n_total) if block_size else n_total
    return torch.tensor(offs, dtype=torch.int64, device=device), len(offs) - 1, max_dim


def kl_naux(t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, omega):
    dev = t_loc.device
    kl = torch.empty(nb, dtype=torch.float32, device=dev)
    na = torch.empty(nb, dtype=torch.int32, device=dev)
    N.check(N.lib().irec_kl_naux(N.ptr(t_loc), N.ptr(t_scale), N.ptr(p_loc), N.ptr(p_scale), N.ptr(gather_idx),
                                 N.ptr(block_offsets), nb, float(omega), N.ptr(kl), N.ptr(na), N.stream_ptr()),
            "irec_kl_naux")
    return kl, na


def schedule(t_loc, t_scale, p_loc, p_scale, omega, max_aux=4096):
    """(kl, n_aux, sa, A, E, M) of one block: the per-partition per-dim coefficients the kernels evaluate on the fly
    (include/irec.h: irec_schedule); sa/A/E/M are [n_aux, D] device tensors.  Diagnostics / tests: reads n_aux back."""
    dev = t_loc.device
    D = t_loc.numel()
    kl = torch.empty(1, dtype=torch.float32, device=dev)
    na = torch.empty(1, dtype=torch.int32, device=dev)
    outs = [torch.zeros((max_aux, D), dtype=torch.float32, device=dev) for _ in range(4)]
    N.check(N.lib().irec_schedule(N.ptr(t_loc), N.ptr(t_scale), N.ptr(p_loc), N.ptr(p_scale), D, float(omega), int(max_aux),
                                  N.ptr(kl), N.ptr(na), *(N.ptr(o) for o in outs), N.stream_ptr()), "irec_schedule")
    n = int(na.cpu()[0])
    if n <= 0 or n > max_aux:
        raise CodingError(f"schedule: KL divergence is not finite, zero, or needs more than {max_aux} auxiliary variables (n_aux={n})")
    return float(kl.cpu()[0]), n, *(o[:n] for o in outs)


def _raise_status(status, n_aux, what):
    if not bool((status != N.BLK_OK).any()):
        return
    bad = (status != N.BLK_OK).nonzero()
    if bad.numel():
        b = int(bad[0])
        st = int(status[b])
        if st == N.BLK_BAD_KL:
            raise CodingError(f"{what}: block {b}: KL divergence is not finite or needs no auxiliary variable "
                              f"(n_aux={int(n_aux[b])})")
        raise CodingError(f"{what}: block {b}: number of auxiliary variables {int(n_aux[b])} exceeds the capacity")


class BeamEncodeResult:
    __slots__ = ("indices", "n_aux", "sample", "kl")

    def __init__(self, indices, n_aux, sample, kl):
        self.indices, self.n_aux, self.sample, self.kl = indices, n_aux, sample, kl


# Row capacity (`max_aux`) of a launch when the caller has not promised one: the last need seen for the same coder
# parameters, with head-room.  The kernels compute KL and n_aux themselves and flag rows that do not fit
# (IREC_BLK_TOO_LONG); only then the launch is repeated with the exact capacity -- no sizing pre-pass, no
# device->host read before the launch.
_AUX_HINT_FIRST = 128
_aux_hint = {}


def _hint_get(key):
    return _aux_hint.get(key, _AUX_HINT_FIRST)


def _hint_update(key, seen_max):
    want = max(_AUX_HINT_FIRST, (int(seen_max * 1.25) + 63) // 64 * 64)
    cur = _aux_hint.get(key)
    if cur is None or want > cur or want < cur // 2:
        _aux_hint[key] = want


def _too_long_need(status, n_aux, max_aux, what):
    """capacity a repeat launch needs, or None when every row fitted; rows flagged although they would fit the
    capacity exceed the auxiliary-ratio table in force (learned ratios) -- the reference's CodingError"""
    import numpy as np
    st = status.numpy() if hasattr(status, "numpy") else np.asarray(status)
    na = n_aux.numpy() if hasattr(n_aux, "numpy") else np.asarray(n_aux)
    long_rows = st == N.BLK_TOO_LONG
    if not long_rows.any():
        return None
    need = int(na[long_rows].max())
    ratio_len = int(N.load_library().irec_aux_ratio_len())
    if need > ratio_len or need <= max_aux:
        b = int(np.flatnonzero(long_rows)[0])
        raise CodingError(f"{what}: block {b}: KL divergence higher than auxiliary variables can account for "
                          f"(needs {int(na[b])}, the ratio table covers {ratio_len})")
    return need


class PendingBeamResult:
    """Result of a launched beam encode whose index lists have not been read back yet: the packed (status, n_aux, indices)
    rows are on their way to pinned host memory; `indices()` waits for that copy only (not for later work on the stream),
    checks the per-block status and builds the Python lists.  Lets a caller that codes several tensors in a row (the
    levels of a model, a batch of images) overlap the host-side list building of one launch with the next launch.
    Without a promised `max_aux` a row may not have fitted the guessed capacity: `indices()` then repeats the launch
    (same output tensors, `retried` = True) -- `sample` is final only once `indices()` has returned."""
    __slots__ = ("_host", "_event", "_nb", "sample", "kl", "_lists", "_retry", "_max_aux", "_hint_key", "retried")

    def __init__(self, host, event, nb, sample, kl, retry=None, max_aux=0, hint_key=None):
        self._host, self._event, self._nb, self.sample, self.kl, self._lists = host, event, nb, sample, kl, None
        self._retry, self._max_aux, self._hint_key, self.retried = retry, max_aux, hint_key, False

    def indices(self):
        if self._lists is None:
            self._event.synchronize()
            packed = self._host
            status, n_aux = packed[:, 0], packed[:, 1]
            if self._retry is not None:
                need = _too_long_need(status, n_aux, self._max_aux, "beam encode")
                if need is not None:
                    packed = self._retry(need)
                    status, n_aux = packed[:, 0], packed[:, 1]
                    self.retried = True
                if self._hint_key is not None and packed.shape[0]:
                    _hint_update(self._hint_key, int(n_aux.max()))
            _raise_status(status, n_aux, "beam encode")
            idx_np, na_np = packed[:, 2:].numpy(), n_aux.numpy().tolist()
            self._lists = [idx_np[b, :na_np[b]].tolist() for b in range(self._nb)]
            self._host = self._retry = None
        return self._lists


def beam_encode_blocks(t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, max_block_dim, omega, S, B, seed,
                       max_aux=None, return_device=False, lazy=False):
    """BeamSearchCoder.encode_block over nb blocks in one launch (beam_search_coder.py:53-122).  max_aux: promised row
    capacity (no block needs more auxiliary variables); None = guessed from earlier calls and grown on demand."""
    lib = N.lib()
    dev = t_loc.device
    promised = max_aux is not None
    hint_key = (str(dev), float(omega))
    if not promised:
        max_aux = _hint_get(hint_key)
    out_sample = torch.empty_like(t_loc)

    def launch(cap):
        ws_bytes = int(lib.irec_beam_encode_workspace_bytes(nb, int(max_block_dim), int(S), int(B), int(cap)))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        out_idx = torch.empty((nb, cap), dtype=torch.int32, device=dev)
        out_na = torch.empty(nb, dtype=torch.int32, device=dev)
        out_st = torch.empty(nb, dtype=torch.int32, device=dev)
        N.check(lib.irec_beam_encode(N.ptr(t_loc), N.ptr(t_scale), N.ptr(p_loc), N.ptr(p_scale), N.ptr(gather_idx),
                                     N.ptr(block_offsets), nb, int(max_block_dim), float(omega), int(S), int(B), int(seed),
                                     N.ptr(out_idx), int(cap), N.ptr(out_na), N.ptr(out_st), N.ptr(out_sample),
                                     N.ptr(ws), ws_bytes, N.stream_ptr()), "irec_beam_encode")
        return out_idx, out_na, out_st

    def pack(outs):
        out_idx, out_na, out_st = outs
        return torch.cat([out_st.view(-1, 1), out_na.view(-1, 1), out_idx], dim=1)

    outs = launch(int(max_aux))
    if return_device:
        return BeamEncodeResult(outs[0], outs[1], out_sample, None), outs[2]
    if lazy:
        packed_dev = pack(outs)
        host = torch.empty(packed_dev.shape, dtype=packed_dev.dtype, pin_memory=True)
        host.copy_(packed_dev, non_blocking=True)
        event = torch.cuda.Event()
        event.record()
        retry = None if promised else (lambda need: pack(launch(need)).cpu())
        return PendingBeamResult(host, event, nb, out_sample, None, retry=retry, max_aux=int(max_aux),
                                 hint_key=None if promised else hint_key)
    packed = pack(outs).cpu()                                          # one D2H copy
    if not promised:
        need = _too_long_need(packed[:, 0], packed[:, 1], int(max_aux), "beam encode")
        if need is not None:
            packed = pack(launch(need)).cpu()
        if nb:
            _hint_update(hint_key, int(packed[:, 1].max()))
    status, n_aux, idx = packed[:, 0], packed[:, 1], packed[:, 2:]
    _raise_status(status, n_aux, "beam encode")
    idx_np, na_np = idx.numpy(), n_aux.numpy().tolist()          # numpy slicing: ~10x cheaper per block than tensor indexing
    indices = [idx_np[b, :na_np[b]].tolist() for b in range(nb)]
    return BeamEncodeResult(indices, n_aux, out_sample, None)


def pack_indices(indices, dtype, device):
    """per-block index lists -> ([nb, max_aux] tensor, [nb] counts, max_aux): one flat conversion of the Python ints, one
    vectorised scatter into the padded rows, ONE host->device copy (counts ride in the first column)"""
    import itertools
    import numpy as np
    nb = len(indices)
    np_dtype = np.int64 if dtype == torch.int64 else np.int32
    lens = np.fromiter((len(i) for i in indices), np.int64, count=nb)
    total = int(lens.sum())
    max_aux = max(1, int(lens.max()) if nb else 1)
    host = np.zeros((nb, max_aux + 1), dtype=np_dtype)
    host[:, 0] = lens
    if total:
        flat = np.fromiter(itertools.chain.from_iterable(indices), np_dtype, count=total)
        rows = np.repeat(np.arange(nb), lens)
        cols = np.arange(total) - np.repeat(np.cumsum(lens) - lens, lens) + 1
        host[rows, cols] = flat
    both = torch.from_numpy(host).to(device)
    idx = both[:, 1:].contiguous()
    n = both[:, 0].to(torch.int32).contiguous()
    return idx, n, max_aux


def _check_index_counts(indices, what):
    """host-side validation of per-block index lists before a decode launch (no device read): the reference raises
    CodingError for a list longer than the ratio table (coder.py:226-231)"""
    longest = max((len(i) for i in indices), default=0)
    ratio_len = int(N.load_library().irec_aux_ratio_len())
    if longest > ratio_len:
        raise CodingError(f"{what}: KL divergence higher than auxiliary variables can account for. Maximum possible number "
                          f"of partitions is {ratio_len}. Requested {longest}")


def beam_decode_blocks(p_loc, p_scale, gather_idx, block_offsets, nb, S, seed, indices, return_status=False):
    """BeamSearchCoder.decode_block over nb blocks (beam_search_coder.py:124-148); `indices` is a list
    of per-block index lists in partition order, or a tuple (idx_tensor, n_aux_tensor, max_aux) of device tensors
    (then pass return_status=True to receive the per-block status tensor: rows the kernel refused decode to NaN)."""
    lib = N.lib()
    dev = p_loc.device
    if isinstance(indices, tuple):
        idx, n_aux, max_aux = indices
    else:
        _check_index_counts(indices, "beam decode")
        idx, n_aux, max_aux = pack_indices(indices, torch.int32, dev)
    out = torch.empty_like(p_loc)
    status = torch.empty(nb, dtype=torch.int32, device=dev) if return_status else None
    N.check(lib.irec_beam_decode(N.ptr(p_loc), N.ptr(p_scale), N.ptr(gather_idx), N.ptr(block_offsets), nb, int(S),
                                 int(seed), N.ptr(idx), int(max_aux), N.ptr(n_aux), N.ptr(out), N.ptr(status),
                                 N.stream_ptr()), "irec_beam_decode")
    return (out, status) if return_status else out


def is_encode_blocks(t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, max_block_dim, omega, S, seed, max_aux=None):
    """GaussianCoder.encode_block with an ImportanceSampler over nb blocks (coder.py:493-559).  The kernel computes KL and
    the number of auxiliary variables itself; results come back in ONE device->host copy."""
    lib = N.lib()
    dev = t_loc.device
    promised = max_aux is not None
    hint_key = (str(dev), float(omega), "is")
    if not promised:
        max_aux = _hint_get(hint_key)
    out_sample = torch.empty_like(t_loc)

    def launch(cap):
        ws_bytes = int(lib.irec_is_encode_workspace_bytes(nb, int(max_block_dim), int(S), int(cap)))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        out_idx = torch.empty((nb, cap), dtype=torch.int64, device=dev)
        out_n = torch.empty(nb, dtype=torch.int32, device=dev)
        out_st = torch.empty(nb, dtype=torch.int32, device=dev)
        N.check(lib.irec_is_encode(N.ptr(t_loc), N.ptr(t_scale), N.ptr(p_loc), N.ptr(p_scale), N.ptr(gather_idx),
                                   N.ptr(block_offsets), nb, int(max_block_dim), float(omega), int(S), int(seed),
                                   N.ptr(out_idx), cap, N.ptr(out_n), N.ptr(out_st), N.ptr(out_sample), N.ptr(ws),
                                   ws_bytes, N.stream_ptr()), "irec_is_encode")
        return torch.cat([out_st.view(-1, 1).to(torch.int64), out_n.view(-1, 1).to(torch.int64), out_idx], dim=1).cpu()

    packed = launch(int(max_aux))
    if not promised:
        need = _too_long_need(packed[:, 0], packed[:, 1], int(max_aux), "importance encode")
        if need is not None:
            packed = launch(need)
        if nb:
            _hint_update(hint_key, int(packed[:, 1].max()))
    st_h, n_h, idx_h = packed[:, 0], packed[:, 1], packed[:, 2:]
    if bool((st_h == N.BLK_BAD_KL).any()):
        raise CodingError("importance encode: KL divergence is not finite")
    _raise_status(st_h, n_h, "importance encode")
    idx_np, n_np = idx_h.numpy(), n_h.numpy().tolist()
    indices = [idx_np[b, :n_np[b]].tolist() for b in range(nb)]
    return indices, out_sample


def is_decode_blocks(p_loc, p_scale, gather_idx, block_offsets, nb, max_block_dim, seed, indices):
    """GaussianCoder.decode_block with an ImportanceSampler over nb blocks (coder.py:561-584): one launch, O(n_aux * D) per
    block, the index lists packed and uploaded once."""
    lib = N.lib()
    dev = p_loc.device
    if any(len(i) < 1 for i in indices):
        raise IndexError("importance decode: empty index list")          # reference: indices[0] on an empty list
    _check_index_counts(indices, "importance decode")
    idx, n_idx, max_aux = pack_indices(indices, torch.int64, dev)
    ws_bytes = int(lib.irec_is_block_workspace_bytes(max_aux))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    out = torch.empty_like(p_loc)
    N.check(lib.irec_is_decode(N.ptr(p_loc), N.ptr(p_scale), N.ptr(gather_idx), N.ptr(block_offsets), nb,
                               int(max_block_dim), int(seed), N.ptr(idx), max_aux, N.ptr(n_idx), N.ptr(out), None, N.ptr(ws),
                               ws_bytes, N.stream_ptr()), "irec_is_decode")
    return out


def is_coded_sample(t_loc, t_scale, p_loc, p_scale, S, seed):
    """encode_gaussian_importance_sample, alpha = inf (importance_sampling.py:9-79) -> (index tensor, sample)"""
    lib = N.lib()
    dev = t_loc.device
    D = t_loc.numel()
    ws_bytes = int(lib.irec_is_workspace_bytes(D))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    out_index = torch.empty(1, dtype=torch.int64, device=dev)
    out_sample = torch.empty(D, dtype=torch.float32, device=dev)
    N.check(lib.irec_is_coded_sample(N.ptr(t_loc), N.ptr(t_scale), N.ptr(p_loc), N.ptr(p_scale), D, int(S), int(seed),
                                     N.ptr(out_index), N.ptr(out_sample), N.ptr(ws), ws_bytes, N.stream_ptr()),
            "irec_is_coded_sample")
    return out_index, out_sample


def is_decode_sample(p_loc, p_scale, index, seed):
    lib = N.lib()
    dev = p_loc.device
    D = p_loc.numel()
    if not isinstance(index, torch.Tensor):
        index = torch.tensor([int(index)], dtype=torch.int64, device=dev)
    index = index.to(device=dev, dtype=torch.int64).reshape(1).contiguous()
    out = torch.empty(D, dtype=torch.float32, device=dev)
    N.check(lib.irec_is_decode_sample(N.ptr(p_loc), N.ptr(p_scale), D, N.ptr(index), int(seed), N.ptr(out),
                                      N.stream_ptr()), "irec_is_decode_sample")
    return out


def beam_uniform_int(q, start, n, device="cuda"):
    out = torch.empty(n, dtype=torch.int32, device=device)
    N.check(N.lib().irec_beam_uniform_int(int(q), int(start), int(n), N.ptr(out), N.stream_ptr()), "irec_beam_uniform_int")
    return out


def is_normal_stream(seed, start, n, device="cuda"):
    out = torch.empty(n, dtype=torch.float32, device=device)
    N.check(N.lib().irec_is_normal_stream(int(seed), int(start), int(n), N.ptr(out), N.stream_ptr()), "irec_is_normal_stream")
    return out


def normal_stream_seeded(global_seed, op_seed, start, n, device="cuda"):
    """z[start : start + n] of tf.random.normal(seed=op_seed) after tf.random.set_seed(global_seed)"""
    out = torch.empty(n, dtype=torch.float32, device=device)
    N.check(N.lib().irec_normal_stream_seeded(int(global_seed), int(op_seed), int(start), int(n), N.ptr(out), N.stream_ptr()),
            "irec_normal_stream_seeded")
    return out


def uniform_int_stream(global_seed, op_seed, lo, hi, start, n, device="cuda"):
    """tf.random.uniform(..., lo, hi, int32, seed=op_seed)[start : start + n] after tf.random.set_seed(global_seed)"""
    out = torch.empty(n, dtype=torch.int32, device=device)
    N.check(N.lib().irec_uniform_int_stream(int(global_seed), int(op_seed), int(lo), int(hi), int(start), int(n), N.ptr(out),
                                            N.stream_ptr()), "irec_uniform_int_stream")
    return out


# ------------------------------------------------------------------------------------------------
# candidate-range sharded beam coder for ONE block (large S; 1..8 GPUs).  Every rank holds a replica
# of the block state; per partition each rank scores its contiguous candidate range, the per-rank
# top-B records are all-gathered (B * 16 bytes per rank) and merged identically everywhere.
# ------------------------------------------------------------------------------------------------
class ShardedBeamBlock:
    def __init__(self, D, S, B, omega, max_aux=1024, group=None, device=None, single=False):
        """single=True: ignore an initialised process group and score the whole candidate range on this GPU (the
        unsharded run a sharded result is compared with)"""
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized() and not single) else None
        self.group = group
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.D, self.S, self.B, self.omega, self.max_aux = int(D), int(S), int(B), float(omega), int(max_aux)
        self.device = torch.device(device if device is not None else "cuda")
        lib = N.lib()
        self.state = torch.empty(int(lib.irec_beam_state_bytes(self.D, self.B, self.max_aux)), dtype=torch.uint8,
                                 device=self.device)
        self.ws_bytes = int(lib.irec_beam_step_workspace_bytes(self.D, self.B))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        # one gather buffer: per rank B records (16 B each) followed by the count in the last record slot
        self.local, self.gathered = SH.alloc_exchange(self.B, self.world, self.device)
        self.s_begin, self.s_end = SH.candidate_range(self.S, self.rank, self.world)
        self._graphs = {}
        self.p2p = None
        if self.world > 1 and os.environ.get("IREC_P2P", "1") != "0":
            self.p2p = self._setup_p2p()

    def _setup_p2p(self):
        """symmetric exchange buffers mapped into every rank (torch symmetric memory: NVLink peer access); None if the
        platform cannot provide them (then the records go through an NCCL all-gather)"""
        try:
            import torch.distributed._symmetric_memory as symm
            grp = self.group if self.group is not None else self.dist.group.WORLD
            words = int(N.lib().irec_p2p_exchange_bytes(self.B, self.world)) // 4
            buf = symm.empty(words, dtype=torch.int32, device=self.device)
            buf.zero_()
            hdl = symm.rendezvous(buf, grp)
            ptrs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=self.device)
            out_rec = torch.zeros(self.world * self.B * SH.RECORD_WORDS, dtype=torch.int32, device=self.device)
            out_cnt = torch.zeros(self.world, dtype=torch.int32, device=self.device)
            torch.cuda.synchronize(self.device)
            self.dist.barrier(group=self.group)
            return {"buf": buf, "hdl": hdl, "ptrs": ptrs, "rec": out_rec, "cnt": out_cnt}
        except Exception as e:              # noqa: BLE001  (no peer access / symmetric memory unavailable)
            if os.environ.get("IREC_P2P") == "1":
                raise
            self.p2p_error = repr(e)
            return None

    def init(self, t_loc, t_scale, p_loc, p_scale, seed):
        lib = N.lib()
        N.check(lib.irec_beam_state_init(N.ptr(self.state), N.ptr(t_loc), N.ptr(t_scale), N.ptr(p_loc), N.ptr(p_scale),
                                         None, 0, self.D, self.omega, self.S, self.B, self.max_aux, int(seed),
                                         N.stream_ptr()), "irec_beam_state_init")
        n_aux, status, kl = C.c_int32(0), C.c_int32(0), C.c_float(0)
        N.check(lib.irec_beam_state_query(N.ptr(self.state), C.byref(n_aux), C.byref(status), C.byref(kl),
                                          N.stream_ptr()), "irec_beam_state_query")
        if status.value != N.BLK_OK:
            raise CodingError(f"sharded beam encode: bad block (status {status.value}, n_aux {n_aux.value}, kl {kl.value})")
        self.n_aux = n_aux.value
        return self.n_aux

    def step(self, t):
        """one partition: local score -> all-gather -> merge + commit"""
        lib = N.lib()
        B = self.B
        rec_ptr = C.c_void_p(self.local.data_ptr())
        cnt_ptr = C.c_void_p(self.local.data_ptr() + B * N.RECORD_BYTES)
        N.check(lib.irec_beam_step_score(N.ptr(self.state), self.D, B, t, self.s_begin, self.s_end, 1, rec_ptr, cnt_ptr,
                                         N.ptr(self.ws), self.ws_bytes, N.stream_ptr()), "irec_beam_step_score")
        if self.p2p is not None:
            recs, counts = self.p2p["rec"], self.p2p["cnt"]
            N.check(lib.irec_p2p_exchange(N.ptr(self.p2p["ptrs"]), self.rank, self.world, B, rec_ptr, N.ptr(recs),
                                          N.ptr(counts), N.stream_ptr()), "irec_p2p_exchange")
        else:
            recs, counts = SH.exchange_records(self.local, self.gathered, B, self.world, self.dist, self.group)
        N.check(lib.irec_beam_step_commit(N.ptr(self.state), self.D, B, t, N.ptr(recs), N.ptr(counts), self.world,
                                          N.ptr(self.ws), self.ws_bytes, N.stream_ptr()), "irec_beam_step_commit")

    def capture_steps(self, n_aux):
        """CUDA graph of the n_aux (score -> all-gather -> merge + commit) steps: ~10 small launches and one NCCL
        collective per auxiliary variable are launch-bound below S ~ 2^22, so they are captured once and replayed.
        Call after init() (the graph reads and advances the block state in place); every rank must capture."""
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.graph(graph, stream=side):
            for t in range(n_aux):
                self.step(t)
        torch.cuda.current_stream(self.device).wait_stream(side)
        return graph

    def encode_graphed(self, t_loc, t_scale, p_loc, p_scale, seed, graphs=None):
        """encode() with the step loop replayed from a CUDA graph (cached per n_aux in `graphs`)"""
        n_aux = self.init(t_loc, t_scale, p_loc, p_scale, seed)
        graphs = self._graphs if graphs is None else graphs
        if n_aux not in graphs:
            graphs[n_aux] = self.capture_steps(n_aux)
        graphs[n_aux].replay()
        return self.finish()

    def fused_available(self):
        """the one-launch path needs the block state to fit shared memory (D <= 256 at 20 beams) and, across ranks, the
        peer-memory exchange buffers"""
        return bool(N.lib().irec_beam_fused_fits(self.D, self.B)) and (self.world == 1 or self.p2p is not None)

    def encode_fused(self, t_loc, t_scale, p_loc, p_scale, seed):
        """encode() as ONE cooperative launch for all auxiliary variables (irec_beam_encode_fused): per variable a single
        all-to-all point inside the kernel instead of seven dependent launches; no host read of n_aux before the launch.
        Falls back to the step loop where the fused kernel does not apply."""
        if not self.fused_available():
            return self.encode(t_loc, t_scale, p_loc, p_scale, seed)
        lib = N.lib()
        N.check(lib.irec_beam_state_init(N.ptr(self.state), N.ptr(t_loc), N.ptr(t_scale), N.ptr(p_loc), N.ptr(p_scale),
                                         None, 0, self.D, self.omega, self.S, self.B, self.max_aux, int(seed),
                                         N.stream_ptr()), "irec_beam_state_init")
        if getattr(self, "_fused_ws", None) is None:
            self._fused_ws = torch.empty(int(lib.irec_beam_fused_workspace_bytes(self.B, self.world)), dtype=torch.uint8,
                                         device=self.device)
        peers = N.ptr(self.p2p["ptrs"]) if self.p2p is not None else None
        N.check(lib.irec_beam_encode_fused(N.ptr(self.state), self.D, self.B, self.s_begin, self.s_end, peers, self.rank,
                                           self.world, N.ptr(self._fused_ws), self._fused_ws.numel(), N.stream_ptr()),
                "irec_beam_encode_fused")
        return self.finish(read_n_aux=True)

    def finish(self, read_n_aux=False):
        lib = N.lib()
        idx = torch.zeros(self.max_aux, dtype=torch.int32, device=self.device)
        meta = torch.zeros(2, dtype=torch.int32, device=self.device)
        sample = torch.empty(self.D, dtype=torch.float32, device=self.device)
        N.check(lib.irec_beam_state_finish(N.ptr(self.state), self.D, None, 0, N.ptr(idx),
                                           C.c_void_p(meta.data_ptr()), C.c_void_p(meta.data_ptr() + 4), N.ptr(sample),
                                           N.stream_ptr()), "irec_beam_state_finish")
        if read_n_aux:                      # n_aux / status were never read on the host: one D2H of (n_aux, status, indices)
            host = torch.cat([meta, idx]).cpu()
            self.n_aux, status = int(host[0]), int(host[1])
            if status != N.BLK_OK:
                raise CodingError(f"sharded beam encode: bad block (status {status}, n_aux {self.n_aux})")
            return host[2:2 + self.n_aux].tolist(), sample
        return idx[:self.n_aux].tolist(), sample

    def encode(self, t_loc, t_scale, p_loc, p_scale, seed):
        n_aux = self.init(t_loc, t_scale, p_loc, p_scale, seed)
        for t in range(n_aux):
            self.step(t)
        return self.finish()
