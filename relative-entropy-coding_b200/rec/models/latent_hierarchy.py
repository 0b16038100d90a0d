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
