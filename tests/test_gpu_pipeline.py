"""End-to-end at BASELINE.json's full sizes (configs[1] and configs[2]) through the callers' flow
(rec.models.LatentHierarchy = the compress/decompress loops of resnet_vae.py:803-860 and large_2_level_vae.py:406-456):
coder.encode per level -> .rec file -> read back -> coder.decode per level.  Size-independent properties: the decoded
latents are bit-identical to the encoder's, every level's prior depends on the previous level's decoded latent, the file
round-trips the index lists, and bits/dim from the index stream equal n_aux * log2(S).  A sample of coder-blocks of every
level is also checked against the CPU oracle on the very inputs the GPU coded."""
import os

import numpy as np
import pytest

import synth  # noqa: F401
from oracle import oracle as O
from oracle import ref_io

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def cuda(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return "cuda:0"


def _oracle_spot_check(ladder, latents, level, coder, block_indices, seed, blocks):
    """oracle encode of a few coder-blocks of `level`, fed with the exact posterior/prior the GPU saw"""
    prior = ladder.prior(level, latents[:level])
    post = ladder.posterior(level, latents[:level])
    tl, ts = post.loc.cpu().numpy().reshape(-1), post.scale.cpu().numpy().reshape(-1)
    pl, ps = prior.loc.cpu().numpy().reshape(-1), prior.scale.cpu().numpy().reshape(-1)
    n, bs = tl.size, coder.block_size
    perm = O.shuffle_perm(n, seed)
    lat = latents[level].cpu().numpy().reshape(-1)
    for b in blocks:
        sel = perm[b * bs:min(n, (b + 1) * bs)]
        ref = O.beam_encode_block(tl[sel], ts[sel], pl[sel], ps[sel], coder.kl_per_partition, coder.n_samples, coder.n_beams, seed)
        assert block_indices[level][b] == ref["indices"].tolist(), (level, b)
        assert np.array_equal(bits(lat[sel]), bits(ref["sample"])), (level, b)


def _run(cuda, tmp_path, shapes, recipe, B, extra, image_shape, spot):
    import torch
    from rec.coding import BeamSearchCoder
    from rec.models import LatentHierarchy, SyntheticLadder
    seed = 42
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=B, extra_samples=extra, block_size=1000)
    ladder = SyntheticLadder(shapes, recipe=recipe, data_seed=5, device=cuda)
    model = LatentHierarchy(ladder)
    path = str(tmp_path / "image.rec")
    block_indices, latents = model.compress(seed, coder, file_path=path, image_shape=image_shape)
    assert [len(t) for t in block_indices] == [-(-int(np.prod(s)) // 1000) for s in shapes]
    # the file holds exactly the index lists; decode from the FILE only
    decoded = model.decompress(coder, file_path=path)
    for a, b in zip(latents, decoded):
        assert a.shape == b.shape and torch.equal(a, b)
    # bits/dim: codelength (nats) of the index stream == n_aux * ln S; the arithmetic-coded file stays close to it
    n_idx = sum(len(b) for t in block_indices for b in t)
    dims = sum(int(np.prod(s)) for s in shapes)
    nats = model.get_codelength(block_indices, coder)
    assert np.isclose(nats, n_idx * np.log(coder.n_samples), rtol=1e-12)
    ideal_bits = nats / np.log(2)
    file_bits = 8 * os.path.getsize(path)
    assert ideal_bits <= file_bits <= 1.05 * ideal_bits + 8 * (28 + 16 * len(shapes)) + 64 * len(shapes) + 16 * sum(len(t) for t in block_indices)
    for level, blocks in spot:
        _oracle_spot_check(ladder, latents, level, coder, block_indices, seed, blocks)
    # a corrupted index derails the decode (the ladder is really sequential)
    bad = [[list(b) for b in t] for t in block_indices]
    bad[0][0][0] = (bad[0][0][0] + 1) % coder.n_samples
    wrong = model.decompress(coder, block_indices=bad, seed=seed)
    # (the synthetic coupling is a contraction, so the perturbation fades after a few levels: check the next one)
    assert not torch.equal(wrong[0], latents[0]) and not torch.equal(wrong[1], latents[1])
    if ref_io.available():          # the reference's own reader accepts our file
        back = ref_io.call([{"op": "rec_read", "path": path}])[0]
        assert back["block_indices"] == block_indices and back["seed"] == seed and back["image_shape"] == list(image_shape)
    return ideal_bits / dims, file_bits / dims


def test_lossy_two_level_kodak_shape(cuda, tmp_path):
    """configs[2]: large_level_2_vae latents for a 768x512 image: [8,12,128] then [32,48,196]; n_beams=10, extra_samples=1."""
    ideal, actual = _run(cuda, tmp_path, [(8, 12, 128), (32, 48, 196)], "c3", 10, 1.0, (512, 768, 3),
                         spot=[(0, [0, 12]), (1, [0, 150, 301])])
    assert 0.1 < ideal < 20 and actual >= ideal


def test_lossless_resnet_vae_cifar_shape(cuda, tmp_path):
    """configs[1]: resnet_vae, 24 latent tensors [16,16,32] for a 32x32 image; n_beams=20, extra_samples=1.2"""
    ideal, actual = _run(cuda, tmp_path, [(16, 16, 32)] * 24, "c2", 20, 1.2, (32, 32, 3),
                         spot=[(0, [0, 8]), (11, [3]), (23, [8])])
    assert 0.1 < ideal < 20 and actual >= ideal


def test_pipelined_compress_equals_level_by_level(cuda, tmp_path):
    """LatentHierarchy.compress(max_aux=...) enqueues the levels without a host synchronisation (coder.encode_lazy): same
    index lists, same latents, same file as the level-by-level loop; a max_aux that is too small is reported."""
    import torch
    from rec.coding import BeamSearchCoder
    from rec.coding.utils import CodingError
    from rec.models import LatentHierarchy, SyntheticLadder
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2, block_size=1000)
    model = LatentHierarchy(SyntheticLadder([(16, 16, 32)] * 5, recipe="c2", data_seed=9, device=cuda))
    a_path, b_path = str(tmp_path / "a.rec"), str(tmp_path / "b.rec")
    idx_a, lat_a = model.compress(42, coder, file_path=a_path)
    idx_b, lat_b = model.compress(42, coder, file_path=b_path, max_aux=256)
    assert idx_a == idx_b and all(torch.equal(x, y) for x, y in zip(lat_a, lat_b))
    assert open(a_path, "rb").read() == open(b_path, "rb").read()
    flat = BeamSearchCoder(kl_per_partition=3., n_beams=10, extra_samples=1.0)          # no block_size: one block per level
    small = LatentHierarchy(SyntheticLadder([(4, 4, 8)] * 3, recipe="c2", data_seed=2, device=cuda))
    i1, l1 = small.compress(7, flat)
    i2, l2 = small.compress(7, flat, max_aux=64)
    assert i1 == i2 and all(torch.equal(x, y) for x, y in zip(l1, l2))
    with pytest.raises(CodingError):
        model.compress(42, coder, max_aux=8)


@pytest.mark.parametrize("block_size", [1000, None])
def test_compress_batch_equals_per_image(cuda, block_size, tmp_path):
    """BASELINE configs[3] as a pipeline with real level-to-level dependence: LatentHierarchy.compress_batch (one launch per
    level and sub-batch, sub-batches on their own CUDA streams, no host synchronisation between levels) == looping `compress`
    over the images, whatever the number of streams; decompress_batch replays it bit for bit"""
    import torch
    from rec.coding import BeamSearchCoder
    from rec.models import BatchedSyntheticLadder, LatentHierarchy, SyntheticLadder
    shapes = [(16, 16, 32)] * 3 if block_size else [(4, 4, 8)] * 3
    n_images = 6
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2, block_size=block_size)
    batched = LatentHierarchy(BatchedSyntheticLadder(shapes, n_images, recipe="c2", data_seed=40, device=cuda))
    ref_idx, ref_lat = [], []
    for i in range(n_images):
        single = LatentHierarchy(SyntheticLadder(shapes, recipe="c2", data_seed=40 + i, device=cuda))
        bi, lat = single.compress(seed=42, coder=coder)
        ref_idx.append(bi)
        ref_lat.append(lat)
    for n_streams in (1, 2, 3, None):
        idx, lat = batched.compress_batch(seed=42, coder=coder, n_streams=n_streams)
        torch.cuda.synchronize()
        assert idx == ref_idx, f"n_streams={n_streams}: index lists differ from the per-image loop"
        for level in range(len(shapes)):
            for i in range(n_images):
                assert torch.equal(lat[level][i], ref_lat[i][level][0]), (n_streams, level, i)
    dec = batched.decompress_batch(coder, idx, seed=42)
    for level in range(len(shapes)):
        assert torch.equal(dec[level], lat[level])
    # one .rec file per image, byte-identical to what the per-image loop writes; decode from the files
    paths = [str(tmp_path / f"img{i}.rec") for i in range(n_images)]
    batched.compress_batch(seed=42, coder=coder, file_paths=paths)
    single0 = LatentHierarchy(SyntheticLadder(shapes, recipe="c2", data_seed=40, device=cuda))
    single0.compress(seed=42, coder=coder, file_path=str(tmp_path / "single0.rec"))
    assert open(paths[0], "rb").read() == open(str(tmp_path / "single0.rec"), "rb").read()
    dec_f = batched.decompress_batch(coder, file_paths=paths)
    for level in range(len(shapes)):
        assert torch.equal(dec_f[level], lat[level])
