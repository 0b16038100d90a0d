"""examples/lossless and examples/lossy end to end with random-init networks: the PyTorch restatements of
the reference's models call the coder exactly as the reference does (resnet_vae.py:470,476,803-860;
large_2_level_vae.py:406-456) and must decode what they encoded -- from the index lists (RVAE) and from the `.rec`
file alone (lossy VAE)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return "cuda:0"


@pytest.mark.parametrize("num_res_blocks", [6, 24])
def test_rvae_compress_decompress(cuda, num_res_blocks):
    """configs[1] shape: 32x32 image, every level a [1,16,16,32] latent, block_size 1000, beam search B=20, 1+eps=1.2;
    24 residual blocks = the configuration BASELINE.json names (resnet_vae.py:803-836), 6 = the quick variant"""
    import torch
    from rec.models import BidirectionalResNetVAE
    torch.manual_seed(0)
    model = BidirectionalResNetVAE(num_res_blocks=num_res_blocks, sampler="beam_search",
                                   sampler_args={"n_beams": 20, "extra_samples": 1.2}, coder_args={"block_size": 1000},
                                   deterministic_filters=64, stochastic_filters=32, kl_per_partition=3.).to(cuda)
    image = (torch.rand(1, 32, 32, 3, device=cuda) - 0.5)
    model(image)                                     # data-dependent initialisation pass (reference: first call)
    model(image)
    block_indices, reconstruction = model.compress(image, seed=42)
    assert len(block_indices) == num_res_blocks and all(len(level) == 9 for level in block_indices)        # 8 x 1000 + 192 dims
    assert all(isinstance(i, int) and 0 <= i < 36 for level in block_indices for blk in level for i in blk)
    enc_latent_priors = [(blk.prior.loc.clone(), blk.prior.scale.clone()) for blk in model.residual_blocks]
    # reference quirk kept: the model hands every level's LIST OF BLOCKS to coder.get_codelength (resnet_vae.py:838-842),
    # whose beam-search version is len(.) * ln S (beam_search_coder.py:150-151) -- blocks, not auxiliary variables
    nats = model.get_codelength(block_indices)
    assert np.isclose(nats, num_res_blocks * 9 * np.log(36))
    per_block = sum(model.residual_blocks[0].coder.get_codelength(blk) for level in block_indices for blk in level)
    assert np.isclose(per_block, sum(len(blk) for level in block_indices for blk in level) * np.log(36))
    decoded = model.decompress([[list(b) for b in level] for level in block_indices], seed=42, height=32, width=32)
    # the decoder recomputes every prior from the latents it decoded: same convolutions on the same values
    for blk, (pl, ps) in zip(model.residual_blocks, enc_latent_priors):
        assert torch.allclose(blk.prior.loc, pl, atol=1e-5) and torch.allclose(blk.prior.scale, ps, atol=1e-5)
    assert torch.allclose(decoded, reconstruction + 0.5, atol=1e-4)
    assert decoded.shape == (1, 32, 32, 3)
    # the blocks enqueued without a host synchronisation (encode_lazy): same code, same reconstruction
    lazy_indices, lazy_reconstruction = model.compress(image, seed=42, max_aux=256)
    assert lazy_indices == block_indices and torch.equal(lazy_reconstruction, reconstruction)


def test_rvae_compress_batch(cuda):
    """BASELINE.json configs[3] with the real (random-init) networks: a batch of images, one coder launch per residual block;
    decompress_batch rebuilds the batch from the index lists; every image's code has the shape `compress` gives"""
    import torch
    from rec.models import BidirectionalResNetVAE
    torch.manual_seed(0)
    model = BidirectionalResNetVAE(num_res_blocks=6, sampler="beam_search", sampler_args={"n_beams": 20, "extra_samples": 1.2},
                                   coder_args={"block_size": 1000}, deterministic_filters=64, stochastic_filters=32,
                                   kl_per_partition=3.).to(cuda)
    images = torch.rand(5, 32, 32, 3, device=cuda) - 0.5
    model(images)
    model(images)
    codes, reconstruction = model.compress_batch(images, seed=42)
    assert len(codes) == 5 and all(len(img) == 6 and all(len(level) == 9 for level in img) for img in codes)
    assert all(isinstance(i, int) and 0 <= i < 36 for img in codes for level in img for blk in level for i in blk)
    assert len({tuple(tuple(b) for level in img for b in level) for img in codes}) == 5          # different images, different codes
    decoded = model.decompress_batch(codes, seed=42, height=32, width=32)
    assert decoded.shape == (5, 32, 32, 3)
    assert torch.allclose(decoded, reconstruction + 0.5, atol=1e-4)


def test_rvae_importance_sampler(cuda):
    """`sampler="importance"` builds GaussianCoder(ImportanceSampler) (resnet_vae.py:127-133)"""
    import torch
    from rec.models import BidirectionalResNetVAE
    torch.manual_seed(1)
    model = BidirectionalResNetVAE(num_res_blocks=2, sampler="importance", sampler_args={"coding_bits": 6},
                                   coder_args={"block_size": 500}, deterministic_filters=32, stochastic_filters=8,
                                   kl_per_partition=3.).to(cuda)
    image = (torch.rand(1, 16, 16, 3, device=cuda) - 0.5)
    model(image)
    model(image)
    block_indices, reconstruction = model.compress(image, seed=7)
    decoded = model.decompress([[list(b) for b in level] for level in block_indices], seed=7, height=16, width=16)
    assert torch.allclose(decoded, reconstruction + 0.5, atol=1e-4)


def test_lossy_two_level_file_round_trip(cuda, tmp_path):
    """compress_with_lossy_model.py flow: compress -> .rec file -> decompress from the file alone"""
    import torch
    from rec.coding import BeamSearchCoder
    from rec.io.utils import read_compressed_code
    from rec.models import Large2LevelVAE
    torch.manual_seed(0)
    model = Large2LevelVAE(level_1_filters=48, level_2_filters=32).to(cuda)
    image = torch.rand(128, 192, 3, device=cuda)
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=10, extra_samples=1., block_size=1000)
    path = str(tmp_path / "kodak_like.rec")
    reconstruction = model.compress(path, image, seed=42, sampler=coder, block_size=1000, max_index=coder.n_samples)
    seed, image_shape, block_size, block_indices = read_compressed_code(file_path=path)
    assert seed == 42 and tuple(image_shape) == (128, 192, 3) and block_size == 1000
    assert [len(level) for level in block_indices] == [1, -(-8 * 12 * 48 // 1000)]       # [1,2,3,32] and [1,8,12,48]
    decoded = model.decompress(path, sampler=coder)
    assert decoded.shape == (1, 128, 192, 3)
    assert torch.allclose(decoded, reconstruction, atol=1e-4)
