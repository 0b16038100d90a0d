"""The compress / decompress loops of the reference's models, without the networks.

Reference call pattern (the part every model shares):
  * lossless RVAE     rec/models/resnet_vae.py:803-836 (compress), :838-842 (get_codelength), :844-860 (decompress):
                      for each residual block, in generative order:  indices, latent = block.coder.encode(posterior,
                      prior, seed=seed);  the next block's prior is computed from `latent`
  * lossy 2-level VAE rec/models/lossy/large_2_level_vae.py:406-456: `sampler.encode(target, coder, seed=seed)` per level,
                      then `write_compressed_code(...)`; decompress = `read_compressed_code` + `sampler.decode` per level

`LatentHierarchy` is that loop over an abstract ladder: `ladder.prior(level, previous_latents)` and
`ladder.posterior(level, previous_latents)` return distributions with `.loc` / `.scale` of shape [1, ...] (the networks'
job in the reference; `SyntheticLadder` is a seeded stand-in whose priors really depend on the previous level's decoded
latent, so a wrong sample on any level derails every later one -- the same sequential dependence the real models have).
"""
import numpy as np
import torch

from irec_b200.distributions import Normal
from rec.io.utils import read_compressed_code, write_compressed_code


class LatentHierarchy:
    def __init__(self, ladder):
        self.ladder = ladder

    # -- resnet_vae.py:803-836 / large_2_level_vae.py:406-419 -------------------------------------------------------
    def compress(self, seed, coder, file_path=None, image_shape=(32, 32, 3), max_index=None, max_aux=None):
        """codes every level in order; returns (block_indices, latents).  With file_path the index stream is written as
        a `.rec` file (rec/io/utils.py:7-106); max_index defaults to the coder's alphabet size minus one.
        max_aux: promise that no coder-block needs more auxiliary variables than this -- the levels are then enqueued back
        to back without a host synchronisation (coder.encode_lazy; the next level's prior only needs the latent ON THE
        DEVICE) and the index lists are read back at the end; a broken promise raises CodingError, call again without it."""
        block_indices, latents = [], []
        pipelined = max_aux is not None and hasattr(coder, "encode_lazy")
        pending = []
        for level in range(self.ladder.n_levels):
            prior = self.ladder.prior(level, latents)
            posterior = self.ladder.posterior(level, latents)
            if pipelined:
                get_indices, latent = coder.encode_lazy(posterior, prior, seed=seed, max_aux=max_aux)
                pending.append(get_indices)
                latents.append(latent)
                continue
            indices, latent = coder.encode(posterior, prior, seed=seed)
            if coder.block_size is None:
                indices = [indices]                       # one coder-block
            block_indices.append([[int(i) for i in blk] for blk in indices])
            latents.append(latent)
        for get_indices in pending:
            indices = get_indices()
            if coder.block_size is None:
                indices = [indices]
            block_indices.append([[int(i) for i in blk] for blk in indices])
        if file_path is not None:
            if max_index is None:
                max_index = int(getattr(coder, "n_samples", 0)) or (1 + max(i for t in block_indices for b in t for i in b))
            write_compressed_code(file_path=file_path, seed=seed, image_shape=tuple(image_shape),
                                  block_size=coder.block_size or 0, block_indices=block_indices, max_index=max_index)
        return block_indices, latents

    # -- resnet_vae.py:838-842 ----------------------------------------------------------------------------------------
    def get_codelength(self, block_indices, coder):
        """nats, summed over levels and coder-blocks (BeamSearchCoder: n_aux * ln S, beam_search_coder.py:150-151)"""
        return float(sum(coder.get_codelength(blk) for level in block_indices for blk in level))

    # -- resnet_vae.py:844-860 / large_2_level_vae.py:421-456 ---------------------------------------------------------
    def decompress(self, coder, file_path=None, block_indices=None, seed=None):
        """replays the ladder from the index stream (a `.rec` file or the nested lists); returns the latents"""
        if file_path is not None:
            seed, _, _, block_indices = read_compressed_code(file_path=file_path)
        latents = []
        for level in range(self.ladder.n_levels):
            prior = self.ladder.prior(level, latents)
            idx = [list(b) for b in block_indices[level]]
            latent = coder.decode(prior, idx if coder.block_size is not None else idx[0], seed=seed)
            latents.append(latent)
        return latents


    # -- a batch of images (BASELINE.json configs[3]) ----------------------------------------------------------------
    RESERVED_SMS = 4      # SMs left to the ladder's small kernels when sub-batches run on several streams

    @staticmethod
    def _sub_batches(n_images, blocks_per_image, n_streams, device):
        """image ranges coded as independent pipelines, each on its own stream: a launch of fewer than ~8 coder-blocks per block
        context (two contexts per SM) is split in two, so that one launch's tail is filled by the next launch's head
        (DESIGN.md section 5).  bench_batch.py, 128 images x 24 dependent levels on one B200: 683 ms with one stream, 615 ms
        with two and 4 SMs reserved for the ladder's kernels (without the reserve: 640-770 ms, unstable -- a persistent coder
        CTA holds the whole register file of its SM, so the other sub-batch's small kernels starve)."""
        if n_streams is None:
            sms = torch.cuda.get_device_properties(device).multi_processor_count
            n_streams = 2 if n_images * blocks_per_image <= 8 * 2 * sms else 1
        k = max(1, min(int(n_streams), n_images))
        return [(i * n_images // k, (i + 1) * n_images // k) for i in range(k)]

    def compress_batch(self, seed, coder, n_streams=None, file_paths=None, image_shape=(32, 32, 3), max_index=None,
                       _again=True, _reserve=True):
        """every image of a batched ladder (`ladder.n_images`; `ladder.prior(level, latents, lo, hi)` with loc/scale of shape
        [hi - lo, ...]) through all levels: what looping `compress` over the images computes, as ONE launch per level and
        sub-batch (`coder.encode_batch(lazy=True)`).  The levels of an image are sequential (its next prior needs its latent),
        the images are not: the batch is cut into `n_streams` sub-batches, each running its own level sequence on its own
        CUDA stream without host synchronisation; the index lists are read back at the end.
        Returns (block_indices[image][level][block], latents[level] of shape [n_images, ...])."""
        n_images, n_levels = self.ladder.n_images, self.ladder.n_levels
        dev = self.ladder.device
        n0 = int(np.prod(self.ladder.shapes[0]))
        per_image = 1 if coder.block_size is None else -(-n0 // coder.block_size)
        subs = self._sub_batches(n_images, per_image, n_streams, dev)
        if len(subs) > 1 and _reserve:
            from irec_b200.native import reserved_sms
            with reserved_sms(self.RESERVED_SMS):
                return self.compress_batch(seed, coder, n_streams=len(subs), file_paths=file_paths, image_shape=image_shape,
                                           max_index=max_index, _again=_again, _reserve=False)
        cur = torch.cuda.current_stream(dev)
        streams = [torch.cuda.Stream(device=dev) for _ in subs] if len(subs) > 1 else [cur]
        fork = torch.cuda.Event()
        fork.record(cur)
        state = [{"latents": []} for _ in subs]
        block_indices = [[None] * n_levels for _ in range(n_images)]
        wrap = coder.block_size is None

        retried = []

        def drain(level, pending):
            """index lists of a level: built on the host while the NEXT level runs on the device"""
            for (lo, hi), get_indices in zip(subs, pending):
                per_item = get_indices()
                retried.append(bool(getattr(get_indices, "retried", False)))
                for i in range(hi - lo):
                    block_indices[lo + i][level] = [per_item[i]] if wrap else per_item[i]

        previous = None
        for level in range(n_levels):
            pending = []
            for (lo, hi), st, stream in zip(subs, state, streams):
                with torch.cuda.stream(stream):
                    if level == 0 and stream is not cur:
                        stream.wait_event(fork)
                    prior = self.ladder.prior(level, st["latents"], lo, hi)
                    posterior = self.ladder.posterior(level, st["latents"], lo, hi)
                    get_indices, latent = coder.encode_batch(posterior, prior, seed=seed, lazy=True)
                    st["latents"].append(latent)
                    pending.append(get_indices)
            if previous is not None:
                drain(level - 1, previous)
            previous = pending
        drain(n_levels - 1, previous)
        for stream in streams:
            if stream is not cur:
                cur.wait_stream(stream)
        if any(retried):
            # a launch ran out of index rows and was repeated AFTER later levels had consumed its latent: the row-capacity
            # hint has grown meanwhile, so one more pass runs without repeats
            if not _again:
                raise RuntimeError("compress_batch: index-row capacity still too small on the second pass")
            torch.cuda.synchronize(dev)
            return self.compress_batch(seed, coder, n_streams=n_streams, file_paths=file_paths, image_shape=image_shape,
                                       max_index=max_index, _again=False, _reserve=_reserve)
        latents = []
        for level in range(n_levels):
            parts = [st["latents"][level] for st in state]
            for t in parts:
                t.record_stream(cur)
            latents.append(torch.cat(parts, dim=0) if len(parts) > 1 else parts[0])
        if file_paths is not None:                 # one `.rec` file per image (rec/io/utils.py:7-106), as `compress` writes it
            if len(file_paths) != n_images:
                raise ValueError("compress_batch: one file path per image")
            if max_index is None:
                max_index = int(getattr(coder, "n_samples", 0)) or (1 + max(i for img in block_indices for t in img for b in t for i in b))
            for path, per_image in zip(file_paths, block_indices):
                write_compressed_code(file_path=path, seed=seed, image_shape=tuple(image_shape), block_size=coder.block_size or 0,
                                      block_indices=per_image, max_index=max_index)
        return block_indices, latents

    def decompress_batch(self, coder, block_indices=None, seed=None, file_paths=None):
        """replays the batched ladder from the index lists of `compress_batch` (or from its `.rec` files, one per image);
        returns latents[level] of shape [n_images, ...]"""
        n_images = self.ladder.n_images
        if file_paths is not None:
            block_indices = []
            for path in file_paths:
                seed_i, _, _, per_image = read_compressed_code(file_path=path)
                if seed is not None and seed_i != seed:
                    raise ValueError("decompress_batch: the files were coded with different seeds")
                seed = seed_i
                block_indices.append(per_image)
        latents = []
        for level in range(self.ladder.n_levels):
            prior = self.ladder.prior(level, latents, 0, n_images)
            idx = [[list(b) for b in block_indices[i][level]] for i in range(n_images)]
            if coder.block_size is None:
                idx = [item[0] for item in idx]
            latents.append(coder.decode_batch(prior, idx, seed=seed))
        return latents


class SyntheticLadder:
    """Seeded synthetic stand-in for the networks (SURVEY.md 8d recipes): level shapes as given; the prior of level k is
    the recipe's prior shifted by a fixed random projection of the previous level's latent, the posterior is the recipe's
    posterior expressed relative to that prior (so KL per dim matches the recipe)."""

    def __init__(self, shapes, recipe="c2", data_seed=0, device="cuda", coupling=0.25):
        self.shapes = [tuple(s) for s in shapes]
        self.n_levels = len(self.shapes)
        self.device = device
        self.coupling = float(coupling)
        self._base = []
        for k, shape in enumerate(self.shapes):
            n = int(np.prod(shape))
            rng = np.random.Generator(np.random.PCG64(data_seed * 1000 + k))
            pl = (0.5 * rng.standard_normal(n)).astype(np.float32)
            if recipe == "c3":
                ps = (np.log1p(np.exp(rng.standard_normal(n))) + 1e-7).astype(np.float32)
            else:
                ps = np.exp(0.3 * rng.standard_normal(n)).astype(np.float32)
            dz = (0.4 * rng.standard_normal(n)).astype(np.float32)           # (mu_t - mu_p) / sigma_p
            rs = np.exp(rng.uniform(-1.0, 0.0, n)).astype(np.float32)        # sigma_t / sigma_p
            mix = rng.integers(0, 1 << 30, n)                                # which dim of the previous latent feeds dim i
            self._base.append([torch.from_numpy(a).to(device) for a in (pl, ps, dz, rs)] + [torch.from_numpy(mix).to(device)])

    def _shift(self, level, latents):
        if level == 0 or not latents:
            return 0.0
        prev = latents[level - 1].reshape(-1)
        mix = self._base[level][4] % prev.numel()
        return self.coupling * torch.tanh(prev[mix])

    def prior(self, level, latents):
        pl, ps = self._base[level][0], self._base[level][1]
        shape = (1,) + self.shapes[level]
        return Normal((pl + self._shift(level, latents)).reshape(shape), ps.reshape(shape))

    def posterior(self, level, latents):
        pl, ps, dz, rs, _ = self._base[level]
        shape = (1,) + self.shapes[level]
        mu_p = pl + self._shift(level, latents)
        return Normal((mu_p + ps * dz).reshape(shape), (ps * rs).reshape(shape))


class BatchedSyntheticLadder:
    """`n_images` independent SyntheticLadders side by side (image i uses data_seed + i): prior / posterior return the rows
    [lo, hi) of the batch, shape [hi - lo, ...]; `latents` are the previous levels' latents of the SAME rows.  Row i equals
    SyntheticLadder(shapes, recipe, data_seed + i) bit for bit."""

    def __init__(self, shapes, n_images, recipe="c2", data_seed=0, device="cuda", coupling=0.25):
        self.shapes = [tuple(s) for s in shapes]
        self.n_levels = len(self.shapes)
        self.n_images = int(n_images)
        self.device = device
        self.coupling = float(coupling)
        singles = [SyntheticLadder(shapes, recipe=recipe, data_seed=data_seed + i, device=device, coupling=coupling)
                   for i in range(self.n_images)]
        self._base = [[torch.stack([s._base[k][j] for s in singles]) for j in range(5)] for k in range(self.n_levels)]

    def _shift(self, level, latents, lo, hi):
        if level == 0 or not latents:
            return 0.0
        prev = latents[level - 1].reshape(hi - lo, -1)
        mix = self._base[level][4][lo:hi] % prev.shape[1]
        return self.coupling * torch.tanh(torch.gather(prev, 1, mix))

    def prior(self, level, latents, lo=0, hi=None):
        hi = self.n_images if hi is None else hi
        pl, ps = self._base[level][0][lo:hi], self._base[level][1][lo:hi]
        shape = (hi - lo,) + self.shapes[level]
        return Normal((pl + self._shift(level, latents, lo, hi)).reshape(shape), ps.reshape(shape))

    def posterior(self, level, latents, lo=0, hi=None):
        hi = self.n_images if hi is None else hi
        pl, ps, dz, rs, _ = (a[lo:hi] for a in self._base[level])
        shape = (hi - lo,) + self.shapes[level]
        mu_p = pl + self._shift(level, latents, lo, hi)
        return Normal((mu_p + ps * dz).reshape(shape), (ps * rs).reshape(shape))
