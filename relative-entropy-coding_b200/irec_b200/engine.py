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


class PendingBeamResult:
    """Result of a launched beam encode whose index lists have not been read back yet: the packed (status, n_aux, indices)
    rows are on their way to pinned host memory; `indices()` waits for that copy only (not for later work on the stream),
    checks the per-block status and builds the Python lists.  Lets a caller that codes several tensors in a row (the
    levels of a model, a batch of images) overlap the host-side list building of one launch with the next launch."""
    __slots__ = ("_host", "_event", "_nb", "sample", "kl", "_lists")

    def __init__(self, host, event, nb, sample, kl):
        self._host, self._event, self._nb, self.sample, self.kl, self._lists = host, event, nb, sample, kl, None

    def indices(self):
        if self._lists is None:
            self._event.synchronize()
            packed = self._host
            status, n_aux, idx = packed[:, 0], packed[:, 1], packed[:, 2:]
            _raise_status(status, n_aux, "beam encode")
            idx_np, na_np = idx.numpy(), n_aux.numpy().tolist()
            self._lists = [idx_np[b, :na_np[b]].tolist() for b in range(self._nb)]
            self._host = None
        return self._lists


def beam_encode_blocks(t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, max_block_dim, omega, S, B, seed,
                       max_aux=None, return_device=False, lazy=False):
    """BeamSearchCoder.encode_block over nb blocks in one launch (beam_search_coder.py:53-122)."""
    lib = N.lib()
    dev = t_loc.device
    kl = None
    if max_aux is None:
        kl, na = kl_naux(t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, omega)
        na_h = na.cpu()
        if bool((na_h <= 0).any()):
            b = int((na_h <= 0).nonzero()[0])
            raise CodingError(f"beam encode: block {b}: KL divergence is not finite or zero (n_aux={int(na_h[b])})")
        max_aux = int(na_h.max())
    ws_bytes = int(lib.irec_beam_encode_workspace_bytes(nb, int(max_block_dim), int(S), int(B), int(max_aux)))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    out_idx = torch.empty((nb, max_aux), dtype=torch.int32, device=dev)
    out_na = torch.empty(nb, dtype=torch.int32, device=dev)
    out_st = torch.empty(nb, dtype=torch.int32, device=dev)
    out_sample = torch.empty_like(t_loc)
    N.check(lib.irec_beam_encode(N.ptr(t_loc), N.ptr(t_scale), N.ptr(p_loc), N.ptr(p_scale), N.ptr(gather_idx),
                                 N.ptr(block_offsets), nb, int(max_block_dim), float(omega), int(S), int(B), int(seed),
                                 N.ptr(out_idx), int(max_aux), N.ptr(out_na), N.ptr(out_st), N.ptr(out_sample),
                                 N.ptr(ws), ws_bytes, N.stream_ptr()), "irec_beam_encode")
    if return_device:
        return BeamEncodeResult(out_idx, out_na, out_sample, kl), out_st
    if lazy:
        packed_dev = torch.cat([out_st.view(-1, 1), out_na.view(-1, 1), out_idx], dim=1)
        host = torch.empty(packed_dev.shape, dtype=packed_dev.dtype, pin_memory=True)
        host.copy_(packed_dev, non_blocking=True)
        event = torch.cuda.Event()
        event.record()
        return PendingBeamResult(host, event, nb, out_sample, kl)
    packed = torch.cat([out_st.view(-1, 1), out_na.view(-1, 1), out_idx], dim=1).cpu()   # one D2H copy
    status, n_aux, idx = packed[:, 0], packed[:, 1], packed[:, 2:]
    _raise_status(status, n_aux, "beam encode")
    idx_np, na_np = idx.numpy(), n_aux.numpy().tolist()          # numpy slicing: ~10x cheaper per block than tensor indexing
    indices = [idx_np[b, :na_np[b]].tolist() for b in range(nb)]
    return BeamEncodeResult(indices, n_aux, out_sample, kl)


def pack_indices(indices, dtype, device):
    nb = len(indices)
    max_aux = max(1, max(len(i) for i in indices))
    import numpy as np
    np_dtype = np.int64 if dtype == torch.int64 else np.int32
    host = np.zeros((nb, max_aux), dtype=np_dtype)
    n = np.zeros(nb, dtype=np.int32)
    for b, ind in enumerate(indices):
        k = len(ind)
        n[b] = k
        if k:
            host[b, :k] = np.asarray([int(v) for v in ind], dtype=np_dtype)
    return torch.from_numpy(host).to(device), torch.from_numpy(n).to(device), max_aux


def beam_decode_blocks(p_loc, p_scale, gather_idx, block_offsets, nb, S, seed, indices):
    """BeamSearchCoder.decode_block over nb blocks (beam_search_coder.py:124-148); `indices` is a list
    of per-block index lists in partition order, or a tuple (idx_tensor, n_aux_tensor, max_aux)."""
    lib = N.lib()
    dev = p_loc.device
    if isinstance(indices, tuple):
        idx, n_aux, max_aux = indices
    else:
        idx, n_aux, max_aux = pack_indices(indices, torch.int32, dev)
    out = torch.empty_like(p_loc)
    N.check(lib.irec_beam_decode(N.ptr(p_loc), N.ptr(p_scale), N.ptr(gather_idx), N.ptr(block_offsets), nb, int(S),
                                 int(seed), N.ptr(idx), int(max_aux), N.ptr(n_aux), N.ptr(out), N.stream_ptr()),
            "irec_beam_decode")
    return out


def is_encode_blocks(t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, max_block_dim, omega, S, seed):
    """GaussianCoder.encode_block with an ImportanceSampler over nb blocks (coder.py:493-559)."""
    lib = N.lib()
    dev = t_loc.device
    _, na = kl_naux(t_loc, t_scale, p_loc, p_scale, gather_idx, block_offsets, nb, omega)
    na_h = na.cpu()
    if bool((na_h < 0).any()):
        raise CodingError("importance encode: KL divergence is not finite")
    max_aux = max(1, int(na_h.max()))
    ws_bytes = int(lib.irec_is_block_workspace_bytes(max_aux))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    out_idx = torch.empty((nb, max_aux), dtype=torch.int64, device=dev)
    out_n = torch.empty(nb, dtype=torch.int32, device=dev)
    out_st = torch.empty(nb, dtype=torch.int32, device=dev)
    out_sample = torch.empty_like(t_loc)
    N.check(lib.irec_is_encode(N.ptr(t_loc), N.ptr(t_scale), N.ptr(p_loc), N.ptr(p_scale), N.ptr(gather_idx),
                               N.ptr(block_offsets), nb, int(max_block_dim), float(omega), int(S), int(seed),
                               N.ptr(out_idx), max_aux, N.ptr(out_n), N.ptr(out_st), N.ptr(out_sample), N.ptr(ws),
                               ws_bytes, N.stream_ptr()), "irec_is_encode")
    st_h, n_h, idx_h = out_st.cpu(), out_n.cpu(), out_idx.cpu()
    _raise_status(st_h, n_h, "importance encode")
    idx_np, n_np = idx_h.numpy(), n_h.numpy().tolist()
    indices = [idx_np[b, :n_np[b]].tolist() for b in range(nb)]
    return indices, out_sample


def is_decode_blocks(p_loc, p_scale, gather_idx, block_offsets, nb, max_block_dim, seed, indices):
    lib = N.lib()
    dev = p_loc.device
    idx, n_idx, max_aux = pack_indices(indices, torch.int64, dev)
    ws_bytes = int(lib.irec_is_block_workspace_bytes(max_aux))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    out = torch.empty_like(p_loc)
    N.check(lib.irec_is_decode(N.ptr(p_loc), N.ptr(p_scale), N.ptr(gather_idx), N.ptr(block_offsets), nb,
                               int(max_block_dim), int(seed), N.ptr(idx), max_aux, N.ptr(n_idx), N.ptr(out), N.ptr(ws),
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
    def __init__(self, D, S, B, omega, max_aux=1024, group=None, device=None):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
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

    def finish(self):
        lib = N.lib()
        idx = torch.zeros(self.max_aux, dtype=torch.int32, device=self.device)
        meta = torch.zeros(2, dtype=torch.int32, device=self.device)
        sample = torch.empty(self.D, dtype=torch.float32, device=self.device)
        N.check(lib.irec_beam_state_finish(N.ptr(self.state), self.D, None, 0, N.ptr(idx),
                                           C.c_void_p(meta.data_ptr()), C.c_void_p(meta.data_ptr() + 4), N.ptr(sample),
                                           N.stream_ptr()), "irec_beam_state_finish")
        return idx[:self.n_aux].tolist(), sample

    def encode(self, t_loc, t_scale, p_loc, p_scale, seed):
        n_aux = self.init(t_loc, t_scale, p_loc, p_scale, seed)
        for t in range(n_aux):
            self.step(t)
        return self.finish()
