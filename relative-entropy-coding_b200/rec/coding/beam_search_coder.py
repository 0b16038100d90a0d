"""rec.coding.beam_search_coder -- BeamSearchCoder with the reference's interface
(reference: rec/coding/beam_search_coder.py:13-151) on the sm_100a kernels of libirec.so.

Per coder-block the whole loop of the reference's `encode_block` (schedule, shared-seed candidate
generation, hash mixing, quantile, log-weights, top-B, beam update; reference :53-122) runs inside
one persistent CTA; `encode()` launches all blocks of a tensor at once.
"""
import numpy as np
import torch

from irec_b200 import engine as E
from rec.coding.coder import GaussianCoder, _dist_tensors
from rec.coding.utils import CodingError


class _LazyIndices:
    """callable returned by encode_batch(lazy=True): reads the index lists back on demand.  `retried` (valid after the call)
    tells a pipelining caller that the launch had to be repeated with more index rows -- the sample tensor it handed to
    later launches before this call returned was not final (engine.PendingBeamResult)."""

    def __init__(self, pend, nest):
        self._pend, self._nest = pend, nest

    def __call__(self):
        return self._nest(self._pend.indices())

    @property
    def retried(self):
        return bool(self._pend.retried)


class BeamSearchCoder(GaussianCoder):

    def __init__(self, kl_per_partition, n_beams, extra_samples=1., extrapolate_auxiliary_ratios=True,
                 name="gaussian_encoder", **kwargs):
        super().__init__(name=name, kl_per_partition=kl_per_partition, sampler=None,
                         extrapolate_auxiliary_ratios=extrapolate_auxiliary_ratios, **kwargs)
        self.n_beams = int(n_beams)
        self.n_samples = int(np.exp(kl_per_partition * extra_samples))       # reference :29 (Omega in nats, truncation)
        self.big_prime = 10007
        if self.n_beams < 1 or self.n_samples < 1:
            raise CodingError("n_beams and the number of samples per partition must be at least 1")

    def _uses_kernels(self):
        return True

    # reference :33-35 (kept for API parity; the kernels keep the running sums incrementally)
    def simple_hash(self, matrix):
        m = torch.as_tensor(matrix, dtype=torch.int32)
        w = torch.arange(69, 69 + m.shape[1], dtype=torch.int32)
        return torch.remainder((m * w).sum(dim=1, dtype=torch.int32), self.big_prime - 1) + 1

    def get_pseudo_random_sample(self, dist, n_samples, index_matrix, seed):
        """reference :37-51: the [n_samples, n_beams', D] candidate tensor  quantile((r * h) mod 10007 / 10007) * scale  with
        r = the seeded uniform-int stream and h = simple_hash(index_matrix).  The kernels never materialise it (they gather
        the same table values inside the scoring loop); kept for API parity and for inspection.  `dist` needs `.scale`
        ([1, D] or [D]); returns a float32 CUDA tensor, bit-identical to what the kernels score."""
        scale = E._f32c(dist.scale, "cuda").reshape(-1)
        D = scale.numel()
        r = E.beam_uniform_int(int(seed), 0, int(n_samples) * D, device=scale.device).to(torch.int64).reshape(int(n_samples), 1, D)
        h = self.simple_hash(index_matrix).to(device=scale.device, dtype=torch.int64).reshape(1, -1, 1)
        k = torch.remainder(r * h, self.big_prime)
        T = self.__dict__.get("_ndtri_dev")
        if T is None or T.device != scale.device:
            from irec_b200 import native as N
            T = self.__dict__["_ndtri_dev"] = N.ndtri_table().to(scale.device)
        out = T[k] * scale.reshape(1, 1, D)
        loc = getattr(dist, "loc", None)
        return out if loc is None else out + E._f32c(loc, "cuda").reshape(1, 1, D)      # quantile(p) = ndtri(p) * scale + loc

    def _encode_flat(self, tl, ts, pl, ps, gather, offsets, nb, max_dim, seed):
        with self._ratios_ctx(tl.device):
            res = E.beam_encode_blocks(tl, ts, pl, ps, gather, offsets, nb, max_dim, self.kl_per_partition, self.n_samples,
                                       self.n_beams, seed)
        return res.indices, res.sample

    def _decode_flat(self, pl, ps, gather, offsets, nb, max_dim, seed, indices):
        with self._ratios_ctx(pl.device):
            return E.beam_decode_blocks(pl, ps, gather, offsets, nb, self.n_samples, seed, indices)

    def encode_block(self, target_dist, coding_dist, seed, update_sampler=False, numpy=True):
        if target_dist.loc.shape[0] != 1:
            raise CodingError("For encoding, batch size must be 1.")
        self._check_ratios()
        tl, ts = _dist_tensors(target_dist)
        pl, ps = _dist_tensors(coding_dist)
        shape = tl.shape
        offsets, nb, max_dim = E.make_block_offsets(tl.numel(), None, tl.device)
        indices, sample = self._encode_flat(tl.reshape(-1), ts.reshape(-1), pl.reshape(-1), ps.reshape(-1), None,
                                            offsets, nb, max_dim, seed)
        return indices[0], sample.reshape(shape)

    def decode_block(self, coding_dist, indices, seed):
        self._check_ratios()
        pl, ps = _dist_tensors(coding_dist)
        shape = pl.shape
        offsets, nb, max_dim = E.make_block_offsets(pl.numel(), None, pl.device)
        indices.reverse()                  # reference :127 reverses the caller's list in place
        sample = self._decode_flat(pl.reshape(-1), ps.reshape(-1), None, offsets, nb, max_dim, seed, [indices[::-1]])
        return sample.reshape(shape)

    def get_codelength(self, indicies):
        return len(indicies) * np.log(self.n_samples)               # reference :150-151

    # -- extension: a batch of independent tensors (e.g. images) in one launch ----------------------
    def encode_batch(self, target_dist, coding_dist, seed, lazy=False):
        """Codes every row of a [N, ...] batch independently with the same coding seed -- what looping the
        reference's `encode` over N single-image batches computes.  Returns (indices[N][n_blocks][n_aux], sample).
        With lazy=True the first element is a callable: the kernel has been launched and the sample tensor is valid in
        stream order, the index lists are read back and built when it is called (so that the caller can launch the
        next tensor first)."""
        self._check_ratios()
        tl, ts = _dist_tensors(target_dist)
        pl, ps = _dist_tensors(coding_dist)
        shape = tl.shape
        n_items = shape[0]
        n = tl[0].numel()
        gather = None
        if self.block_size is not None:
            perm = self._permutation(n, seed, tl.device)
            gather = (perm[None, :] + torch.arange(n_items, device=tl.device)[:, None] * n).reshape(-1).contiguous()
        offsets, nb, max_dim = E.make_block_offsets(n, self.block_size, tl.device, n_items=n_items)
        per = nb // n_items
        block_size = self.block_size

        def nest(indices):
            if block_size is None:
                return [indices[i] for i in range(n_items)]
            return [indices[i * per:(i + 1) * per] for i in range(n_items)]

        if lazy:
            with self._ratios_ctx(tl.device):
                pend = E.beam_encode_blocks(tl.reshape(-1), ts.reshape(-1), pl.reshape(-1), ps.reshape(-1), gather, offsets,
                                            nb, max_dim, self.kl_per_partition, self.n_samples, self.n_beams, seed, lazy=True)
            return _LazyIndices(pend, nest), pend.sample.reshape(shape)
        indices, sample = self._encode_flat(tl.reshape(-1), ts.reshape(-1), pl.reshape(-1), ps.reshape(-1), gather,
                                            offsets, nb, max_dim, seed)
        return nest(indices), sample.reshape(shape)

    def encode_lazy(self, target_dist, coding_dist, seed, max_aux=256):
        """`encode` without a host synchronisation: the caller promises that no coder-block needs more than `max_aux`
        auxiliary variables (KL <= max_aux * kl_per_partition per block), so the sizing pre-pass and its device->host
        read are skipped, the kernel is launched and (get_indices, sample) returned at once; `sample` is valid in stream
        order, `get_indices()` reads the index lists back and raises CodingError if the promise did not hold (then the
        sample was NOT valid and everything computed from it must be redone through `encode`).  Lets the sequential levels
        of a model (resnet_vae.py:821-826) be enqueued back to back."""
        self._check_ratios()
        tl, ts = _dist_tensors(target_dist)
        pl, ps = _dist_tensors(coding_dist)
        shape = tl.shape
        for t in (ts, pl, ps):
            if t.shape != shape:
                raise CodingError("All tensor arguments supplied to split must have the same batch dimensions!")
        if shape[0] != 1:
            raise CodingError("For encoding, batch size must be 1.")
        n = tl.numel()
        perm = self._permutation(n, seed, tl.device) if self.block_size is not None else None
        offsets, nb, max_dim = self._block_structure(n, tl.device)
        with self._ratios_ctx(tl.device):
            pend = E.beam_encode_blocks(tl.reshape(-1), ts.reshape(-1), pl.reshape(-1), ps.reshape(-1), perm, offsets, nb,
                                        max_dim, self.kl_per_partition, self.n_samples, self.n_beams, seed,
                                        max_aux=int(max_aux), lazy=True)
        block_size = self.block_size
        return (lambda: pend.indices() if block_size is not None else pend.indices()[0]), pend.sample.reshape(shape)

    def _block_structure(self, n, device):
        """block offsets of a tensor of n dims (cached: the same shapes recur for every level and image)"""
        key = (int(n), self.block_size, str(device))
        cache = self.__dict__.setdefault("_offsets_cache", {})
        if key not in cache:
            if len(cache) > 64:
                cache.clear()
            cache[key] = E.make_block_offsets(n, self.block_size, device)
        return cache[key]

    def decode_batch(self, coding_dist, indices, seed):
        self._check_ratios()
        pl, ps = _dist_tensors(coding_dist)
        shape = pl.shape
        n_items = shape[0]
        n = pl[0].numel()
        gather = None
        if self.block_size is not None:
            perm = self._permutation(n, seed, pl.device)
            gather = (perm[None, :] + torch.arange(n_items, device=pl.device)[:, None] * n).reshape(-1).contiguous()
            flat_idx = [blk for item in indices for blk in item]
        else:
            flat_idx = list(indices)
        offsets, nb, max_dim = E.make_block_offsets(n, self.block_size, pl.device, n_items=n_items)
        sample = self._decode_flat(pl.reshape(-1), ps.reshape(-1), gather, offsets, nb, max_dim, seed, flat_idx)
        return sample.reshape(shape)
