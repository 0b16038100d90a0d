"""rec.coding.coder -- Coder / GaussianCoder with the reference's interface
(reference: rec/coding/coder.py), running on libirec.so.

What is kept: constructor keywords, `encode/decode/encode_block/decode_block/split/merge/
get_codelength`, the duck-typed distributions (`.loc`, `.scale`, shape [1, ...]), the index list
layout, `CodingError` conditions, and the in-place reversal of the caller's index list in
`decode_block` (reference coder.py:564).  What differs by design: everything is stateless and
counter-based (no global RNG is touched), whole tensors are coded by ONE kernel launch over all
coder-blocks instead of a Python loop, and the `print`s of the reference are gone.

Learned auxiliary ratios (`extrapolate_auxiliary_ratios=False`, reference coder.py:197-410) are
calibrated by `update_auxiliary_variance_ratios` (an offline SGD fit, torch autograd on the tensors' device) and then
handed to the kernels as a device table (include/irec.h: irec_set_thread_aux_ratios) in place of the power law.
The rejection sampler (rec/coding/rejection_sampling.py, samplers.py) is host logic over GPU-regenerated candidate streams; its
two unseeded draws in the reference are seeded here.
"""
import abc

import numpy as np
import torch

from irec_b200 import engine as E
from irec_b200 import native as N
from irec_b200.distributions import Normal
from rec.coding.utils import CodingError
from rec.coding.samplers import Sampler, ImportanceSampler

AUX_RATIO_POWER_LAW = -0.7864636765648174     # reference coder.py:16

_DEVICE = "cuda"


def _dist_tensors(dist, device=_DEVICE):
    """(.loc, .scale) of a duck-typed distribution as float32 CUDA tensors"""
    return E._f32c(dist.loc, device), E._f32c(dist.scale, device)


# -- closed-form Gaussian algebra of the reference (coder.py:141-171), on torch tensors; the CUDA
#    kernels evaluate the same float32 sequences on the device, these exist for API parity and for
#    user-written samplers driven through the generic GaussianCoder loop.
def _pow2(x):
    return x * x


def get_auxiliary_coder(coder, auxiliary_var):
    return Normal(torch.zeros_like(coder.loc), torch.sqrt(auxiliary_var))


def get_auxiliary_target(target, coder, auxiliary_var):
    coder_var = _pow2(coder.scale)
    target_var = _pow2(target.scale)
    mean = (target.loc - coder.loc) * auxiliary_var / coder_var
    var = target_var * _pow2(auxiliary_var) / _pow2(coder_var) + auxiliary_var * (coder_var - auxiliary_var) / coder_var
    return Normal(mean, torch.sqrt(var))


def get_conditional_coder(coder, auxiliary_var, auxiliary_sample):
    return Normal(coder.loc + auxiliary_sample, torch.sqrt(_pow2(coder.scale) - auxiliary_var))


def get_conditional_target(target, coder, auxiliary_var, auxiliary_sample):
    coder_var = _pow2(coder.scale)
    target_var = _pow2(target.scale)
    den = target_var * auxiliary_var + coder_var * (coder_var - auxiliary_var)
    mean = coder.loc + (auxiliary_sample * target_var * coder_var
                        + (target.loc - coder.loc) * (coder_var - auxiliary_var) * coder_var) / den
    var = target_var * coder_var * (coder_var - auxiliary_var) / (auxiliary_var * target_var
                                                                    + coder_var * (coder_var - auxiliary_var))
    return Normal(mean, torch.sqrt(var))


class Coder(abc.ABC):
    """reference: rec/coding/coder.py:27-138"""

    def __init__(self, block_size=None, name="encoder", **kwargs):
        self.name = name
        self.block_size = block_size
        self._perm_cache = {}

    # -- block structure ------------------------------------------------------------------------
    def _permutation(self, num_dims, seed, device):
        """tf.random.set_seed(seed); tf.random.shuffle(range(n)) restated in libirec.so (host C++)"""
        key = (int(num_dims), int(seed), str(device))
        if key not in self._perm_cache:
            if len(self._perm_cache) > 64:
                self._perm_cache.clear()
            self._perm_cache[key] = N.split_permutation(num_dims, seed).to(device)
        return self._perm_cache[key]

    def split(self, *args, seed=42):
        """Splits the arguments into conformal blocks (reference coder.py:38-85)"""
        if self.block_size is None:
            raise CodingError("split needs a block_size")
        tensors = [a if isinstance(a, torch.Tensor) else E._f32c(a, _DEVICE) for a in args]
        shape = tensors[0].shape
        for t in tensors:
            if t.shape != shape:
                raise CodingError("All tensor arguments supplied to split must have the same batch dimensions!")
        flat = [t.reshape(-1) for t in tensors]
        num_dims = flat[0].shape[0]
        perm = self._permutation(num_dims, seed, flat[0].device)
        flat = [f[perm] for f in flat]
        return [[f[i:min(i + self.block_size, num_dims)] for i in range(0, num_dims, self.block_size)] for f in flat]

    def merge(self, *args, shape=None, seed=42):
        """Inverse operation to split (reference coder.py:87-122)"""
        if shape is None:
            raise CodingError("Shape cannot be None!")
        tensors = [torch.cat(list(blocks), dim=0) for blocks in args]
        num_dims = tensors[0].shape[0]
        for t in tensors:
            if t.dim() != 1:
                raise CodingError("All supplied tensors to merge must be rank 1!")
            if t.shape[0] != num_dims:
                raise CodingError("All tensors must have the same number of dimensions!")
        perm = self._permutation(num_dims, seed, tensors[0].device)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(num_dims, device=perm.device)
        return [t[inv].reshape(tuple(shape)) for t in tensors]

    @abc.abstractmethod
    def encode(self, target_dist, coding_dist, seed, **kwargs):
        pass

    @abc.abstractmethod
    def decode(self, coding_dist, indices, seed, **kwargs):
        pass

    @abc.abstractmethod
    def encode_block(self, target_dist, coding_dist, seed, **kwargs):
        pass

    @abc.abstractmethod
    def decode_block(self, coding_dist, indices, seed, **kwargs):
        pass


class GaussianCoder(Coder):
    """reference: rec/coding/coder.py:174-587"""

    def __init__(self, kl_per_partition, sampler: Sampler, extrapolate_auxiliary_ratios=True, block_size=None,
                 name="gaussian_encoder", **kwargs):
        super().__init__(name=name, block_size=block_size, **kwargs)
        self.sampler = sampler
        self.kl_per_partition = float(np.float32(kl_per_partition))      # reference :192 casts to float32
        self.extrapolate_auxiliary_ratios = extrapolate_auxiliary_ratios
        self._initialized = False
        if not self.extrapolate_auxiliary_ratios:                        # reference :197-216
            self.aux_variable_variance_ratios = np.array([1.], dtype=np.float32)
            self.average_counts = np.array([1.], dtype=np.float32)
        self._ratio_dev = {}                                             # device -> (version, table)
        self._ratio_version = 0

    # -- ratios -----------------------------------------------------------------------------------
    def get_auxiliary_ratio(self, index):
        """reference :218-231"""
        if self.extrapolate_auxiliary_ratios:
            return np.power(index + 1., AUX_RATIO_POWER_LAW)
        if not self._initialized:
            raise CodingError("Coder has not been initialized yet, please call"
                              "update_auxiliary_variance_ratios() first"
                              " or use extrapolation")
        if index >= self.aux_variable_variance_ratios.shape[0]:
            raise CodingError("KL divergence higher than auxiliary variables can account for. "
                              "Update auxiliary variable ratios with high-enough KL divergence."
                              "Maximum possible number of partitions is {}."
                              "Requested {}".format(self.aux_variable_variance_ratios.shape[0], index + 1))
        return self.aux_variable_variance_ratios[index]

    def _check_ratios(self):
        if not self.extrapolate_auxiliary_ratios:
            self.get_auxiliary_ratio(0)

    def _ratios_ctx(self, device):
        """context in which the kernels read this coder's ratio table (None = the library's power law)"""
        if self.extrapolate_auxiliary_ratios:
            return N.thread_aux_ratios(None)
        self._check_ratios()
        key = str(device)
        cached = self._ratio_dev.get(key)
        if cached is None or cached[0] != self._ratio_version:
            cached = (self._ratio_version, torch.from_numpy(np.ascontiguousarray(self.aux_variable_variance_ratios,
                                                                                 dtype=np.float32)).to(device))
            self._ratio_dev[key] = cached
        return N.thread_aux_ratios(cached[1])

    def update_auxiliary_variance_ratios(self, target_dist, coding_dist, seed=42, **kwargs):
        """reference :233-270: one calibration step from a (target, coder) pair; with block_size the tensors are split
        with the coder's own permutation and every full block is a batch element (the last, short block is left out)"""
        if self.extrapolate_auxiliary_ratios:
            raise CodingError("update_auxiliary_variance_ratios needs extrapolate_auxiliary_ratios=False")
        tl, ts = (torch.as_tensor(t, dtype=torch.float32) for t in (target_dist.loc, target_dist.scale))
        pl, ps = (torch.as_tensor(t, dtype=torch.float32) for t in (coding_dist.loc, coding_dist.scale))
        if self.block_size is None:
            return self.update_block_auxiliary_variance_ratios(Normal(tl, ts), Normal(pl, ps), seed=seed, **kwargs)
        n = tl.numel()
        perm = N.split_permutation(n, seed).to(tl.device)
        nfull = n // self.block_size if n % self.block_size else n // self.block_size - 1     # reference drops [-1] always
        if nfull < 1:
            raise CodingError("update_auxiliary_variance_ratios: need at least two blocks (the last one is left out)")
        take = perm[:nfull * self.block_size]
        stack = lambda t: t.reshape(-1)[take].reshape(nfull, self.block_size)                  # noqa: E731
        return self.update_block_auxiliary_variance_ratios(Normal(stack(tl), stack(ts)), Normal(stack(pl), stack(ps)),
                                                           seed=seed, **kwargs)

    def update_block_auxiliary_variance_ratios(self, target_dist, coding_dist, relative_tolerance=1e-4, max_iters=10000,
                                               learning_rate=0.001, seed=42):
        """reference :272-410.  For ratio = max .. 2 (number of auxiliary variables still to go): SGD on one scalar
        (sigmoid-reparametrised variance ratio) so that the auxiliary variable carries at most Omega nats and leaves at
        most Omega * (ratio - 1); running average over calls; then condition target and coder on a draw of the auxiliary
        variable and continue.  The reference draws that sample from TF's global RNG; here a torch.Generator seeded with
        `seed` (no global state is touched)."""
        if self.extrapolate_auxiliary_ratios:
            raise CodingError("update_block_auxiliary_variance_ratios needs extrapolate_auxiliary_ratios=False")
        t_loc, t_scale = (torch.as_tensor(t, dtype=torch.float32).detach().clone() for t in (target_dist.loc, target_dist.scale))
        c_loc, c_scale = (torch.as_tensor(t, dtype=torch.float32).detach().clone() for t in (coding_dist.loc, coding_dist.scale))
        dev = t_loc.device
        dims = tuple(range(1, t_loc.dim()))
        omega = self.kl_per_partition
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(seed))

        def kl(ql, qs, pl_, ps_):
            dl = torch.log(qs) - torch.log(ps_)
            return (0.5 * ((ql - pl_) / ps_) ** 2 + 0.5 * torch.expm1(2. * dl) - dl).sum(dim=dims)

        total_kl = kl(t_loc, t_scale, c_loc, c_scale)
        num_aux = 1 + torch.floor(total_kl / omega).to(torch.int32)
        max_num = int(num_aux.max())
        cur = self.aux_variable_variance_ratios.shape[0]
        if max_num > cur:
            grown = np.zeros(max_num, dtype=np.float32); grown[:cur] = self.aux_variable_variance_ratios
            counts = np.zeros(max_num, dtype=np.float32); counts[:cur] = self.average_counts
            self.aux_variable_variance_ratios, self.average_counts = grown, counts
        ratios, counts = self.aux_variable_variance_ratios, self.average_counts
        for ratio in range(max_num, 1, -1):
            sel = (num_aux >= ratio).nonzero().reshape(-1)
            n_el = int(sel.numel())
            tl, ts, cl, cs = t_loc[sel], t_scale[sel], c_loc[sel], c_scale[sel]
            tot = kl(tl, ts, cl, cs)
            if ratios[ratio - 1] > 0.:
                init = float(ratios[ratio - 1])
            elif ratio < max_num:
                init = float(ratios[ratio])
            else:
                init = 1. / ratio
            init = min(max(init, 1e-10), 1. - 1e-10)                     # sigmoid_inverse, reference :19-25
            param = torch.tensor(np.log(init) - np.log1p(-init), dtype=torch.float32, device=dev, requires_grad=True)
            prev_loss = np.inf
            tgt, cod = Normal(tl, ts), Normal(cl, cs)
            for _ in range(int(max_iters)):
                r = torch.sigmoid(param)
                aux_var = r * cs * cs
                aux_t = get_auxiliary_target(tgt, cod, aux_var)
                aux_c = get_auxiliary_coder(cod, aux_var)
                aux_kl = kl(aux_t.loc, aux_t.scale, aux_c.loc, aux_c.scale)
                rest = tot - aux_kl
                loss = (torch.where(aux_kl > omega, (aux_kl - omega) ** 2, torch.zeros_like(aux_kl)) +
                        torch.where(rest > omega * (ratio - 1), (rest - omega * (ratio - 1)) ** 2,
                                    torch.zeros_like(rest))).mean()
                grad, = torch.autograd.grad(loss, param)
                with torch.no_grad():
                    param -= learning_rate * grad
                loss_v = float(loss.detach())
                if abs(prev_loss - loss_v) < relative_tolerance:
                    break
                prev_loss = loss_v
            with torch.no_grad():
                r_fit = float(r.detach())
                ratios[ratio - 1] = (ratios[ratio - 1] * counts[ratio - 1] + r_fit * n_el) / (counts[ratio - 1] + n_el)
                counts[ratio - 1] += n_el
                aux_var = float(ratios[ratio - 1]) * cs * cs
                aux_t = Normal(aux_t.loc.detach(), aux_t.scale.detach())     # of the last SGD iterate, as the reference
                sample = aux_t.loc + aux_t.scale * torch.randn(aux_t.loc.shape, generator=gen, device=dev)
                new_t = get_conditional_target(tgt, cod, aux_var, sample)
                new_c = get_conditional_coder(cod, aux_var, sample)
                t_loc[sel], t_scale[sel] = new_t.loc, new_t.scale
                c_loc[sel], c_scale[sel] = new_c.loc, new_c.scale
        self._initialized = True
        self._ratio_version += 1

    # -- sampler dispatch -------------------------------------------------------------------------
    def _fused_importance(self):
        return isinstance(self.sampler, ImportanceSampler) and np.isinf(self.sampler.alpha)

    def _encode_flat(self, tl, ts, pl, ps, gather, offsets, nb, max_dim, seed):
        with self._ratios_ctx(tl.device):
            indices, sample = E.is_encode_blocks(tl, ts, pl, ps, gather, offsets, nb, max_dim, self.kl_per_partition,
                                                 self.sampler.n_samples, seed)
        return indices, sample

    def _decode_flat(self, pl, ps, gather, offsets, nb, max_dim, seed, indices):
        with self._ratios_ctx(pl.device):
            return E.is_decode_blocks(pl, ps, gather, offsets, nb, max_dim, seed, indices)

    # -- whole tensors ----------------------------------------------------------------------------
    def encode(self, target_dist, coding_dist, seed, **kwargs):
        if self.block_size is None:
            return self.encode_block(target_dist, coding_dist, seed, **kwargs)
        if not self._uses_kernels():
            return self._encode_python_blocks(target_dist, coding_dist, seed, **kwargs)
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
        perm = self._permutation(n, seed, tl.device)
        offsets, nb, max_dim = E.make_block_offsets(n, self.block_size, tl.device)
        indices, sample = self._encode_flat(tl.reshape(-1), ts.reshape(-1), pl.reshape(-1), ps.reshape(-1), perm,
                                            offsets, nb, max_dim, seed)
        return indices, sample.reshape(shape)

    def decode(self, coding_dist, indices, seed, **kwargs):
        if self.block_size is None:
            return self.decode_block(coding_dist, indices, seed, **kwargs)
        if not self._uses_kernels():
            return self._decode_python_blocks(coding_dist, indices, seed, **kwargs)
        self._check_ratios()
        pl, ps = _dist_tensors(coding_dist)
        shape = pl.shape
        n = pl.numel()
        perm = self._permutation(n, seed, pl.device)
        offsets, nb, max_dim = E.make_block_offsets(n, self.block_size, pl.device)
        if len(indices) != nb:
            raise CodingError(f"decode expects {nb} index lists (one per block), got {len(indices)}")
        for ind in indices:
            ind.reverse()                  # the reference reverses every block's list in place (coder.py:564)
        try:
            fwd = [ind[::-1] for ind in indices]
            sample = self._decode_flat(pl.reshape(-1), ps.reshape(-1), perm, offsets, nb, max_dim, seed, fwd)
        except Exception:
            for ind in indices:
                ind.reverse()
            raise
        return sample.reshape(shape)

    def _uses_kernels(self):
        return self._fused_importance()

    # -- one block --------------------------------------------------------------------------------
    def encode_block(self, target_dist, coding_dist, seed, update_sampler=False, verbose=False, numpy=True):
        if target_dist.loc.shape[0] != 1:
            raise CodingError("For encoding, batch size must be 1.")
        if update_sampler or not self._uses_kernels():
            return self._encode_block_python(target_dist, coding_dist, seed, update_sampler=update_sampler,
                                             verbose=verbose, numpy=numpy)
        self._check_ratios()
        tl, ts = _dist_tensors(target_dist)
        pl, ps = _dist_tensors(coding_dist)
        shape = tl.shape
        n = tl.numel()
        offsets, nb, max_dim = E.make_block_offsets(n, None, tl.device)
        indices, sample = self._encode_flat(tl.reshape(-1), ts.reshape(-1), pl.reshape(-1), ps.reshape(-1), None,
                                            offsets, nb, max_dim, seed)
        return indices[0], sample.reshape(shape)

    def decode_block(self, coding_dist, indices, seed, **kwargs):
        if not self._uses_kernels():
            return self._decode_block_python(coding_dist, indices, seed)
        self._check_ratios()
        pl, ps = _dist_tensors(coding_dist)
        shape = pl.shape
        offsets, nb, max_dim = E.make_block_offsets(pl.numel(), None, pl.device)
        indices.reverse()                  # reference coder.py:564 (in place)
        sample = self._decode_flat(pl.reshape(-1), ps.reshape(-1), None, offsets, nb, max_dim, seed, [indices[::-1]])
        return sample.reshape(shape)

    def get_codelength(self, indicies):
        return sum([self.sampler.get_codelength(i) for i in indicies])

    # -- generic (plug-in sampler) path: the reference's Python loop over auxiliary variables --------
    def _encode_block_python(self, target_dist, coding_dist, seed, update_sampler=False, verbose=False, numpy=True):
        """reference coder.py:493-559 for arbitrary `Sampler` plug-ins (the partition work happens in
        sampler.coded_sample; ImportanceSampler runs it on the GPU)."""
        target = Normal(*_dist_tensors(target_dist))
        coder = Normal(*_dist_tensors(coding_dist))
        tl, ts, pl, ps = (t.reshape(-1) for t in (target.loc, target.scale, coder.loc, coder.scale))
        offsets, nb, _ = E.make_block_offsets(tl.numel(), None, tl.device)
        _, na = E.kl_naux(tl, ts, pl, ps, None, offsets, 1, self.kl_per_partition)
        num_aux = int(na.cpu()[0])
        if num_aux < 0:
            raise CodingError("KL divergence is not finite")
        indices = []
        gen = torch.Generator(device=target.loc.device)
        gen.manual_seed(int(seed))
        for i in range(num_aux - 1, 0, -1):
            ratio = torch.tensor(np.float32(self.get_auxiliary_ratio(i)), device=coder.loc.device)
            aux_var = ratio * _pow2(coder.scale)
            aux_target = get_auxiliary_target(target, coder, aux_var)
            aux_coder = get_auxiliary_coder(coder, aux_var)
            if update_sampler:
                self.sampler.update(aux_target, aux_coder)
                aux_sample = aux_target.loc + aux_target.scale * torch.randn(aux_target.loc.shape, generator=gen,
                                                                             device=aux_target.loc.device)
            else:
                index, aux_sample = self.sampler.coded_sample(target=aux_target, coder=aux_coder, seed=seed)
                indices.append(index.numpy() if (numpy and hasattr(index, "numpy")) else index)
            seed += 1
            target = get_conditional_target(target, coder, aux_var, aux_sample)
            coder = get_conditional_coder(coder, aux_var, aux_sample)
        if update_sampler:
            self.sampler.update(target, coder)
            sample = target.loc + target.scale * torch.randn(target.loc.shape, generator=gen, device=target.loc.device)
        else:
            index, sample = self.sampler.coded_sample(target=target, coder=coder, seed=seed)
            indices.append(index.numpy() if (numpy and hasattr(index, "numpy")) else index)
        return indices, sample

    def _decode_block_python(self, coding_dist, indices, seed):
        """reference coder.py:561-584"""
        coder = Normal(*_dist_tensors(coding_dist))
        num_aux = len(indices)
        indices.reverse()
        for i in range(num_aux - 1, 0, -1):
            ratio = torch.tensor(np.float32(self.get_auxiliary_ratio(i)), device=coder.loc.device)
            aux_var = ratio * _pow2(coder.scale)
            aux_coder = get_auxiliary_coder(coder, aux_var)
            aux_sample = self.sampler.decode_sample(coder=aux_coder, sample_index=indices[i], seed=seed)
            seed += 1
            coder = get_conditional_coder(coder, aux_var, aux_sample)
        return self.sampler.decode_sample(coder=coder, sample_index=indices[0], seed=seed)

    def _encode_python_blocks(self, target_dist, coding_dist, seed, **kwargs):
        """reference coder.py:412-457 (per-block Python loop) for plug-in samplers"""
        tl, ts = _dist_tensors(target_dist)
        pl, ps = _dist_tensors(coding_dist)
        shape = tl.shape
        blocks = self.split(tl, ts, pl, ps, seed=seed)
        samples, indices = [], []
        for btl, bts, bpl, bps in zip(*blocks):
            ind, samp = self.encode_block(Normal(btl[None, :], bts[None, :]), Normal(bpl[None, :], bps[None, :]), seed,
                                          **kwargs)
            samples.append(samp.reshape(-1))
            indices.append(ind)
        sample, = self.merge(samples, shape=shape, seed=seed)
        return indices, sample

    def _decode_python_blocks(self, coding_dist, indices, seed, **kwargs):
        pl, ps = _dist_tensors(coding_dist)
        shape = pl.shape
        locs, scales = self.split(pl, ps, seed=seed)
        samples = []
        for inds, bpl, bps in zip(indices, locs, scales):
            samp = self.decode_block(Normal(bpl[None, :], bps[None, :]), inds, seed, **kwargs)
            samples.append(samp.reshape(-1))
        sample, = self.merge(samples, shape=shape, seed=seed)
        return sample
