"""The callers' networks on the CPU: layer output sizes and the latent shapes BASELINE.json's configs
quote -- [1,16,16,32] per RVAE level for a 32x32 image (resnet_vae.py:535-539,633), [1,H/64,W/64,l2] and [1,H/16,W/16,l1]
for the two-level lossy VAE (large_2_level_vae.py:26-69,147-179,313).  No coder call here (that needs the GPU)."""
import pytest
import torch

from rec.models.lossy import GDN, Large2LevelVAE, SignalConv2D
from rec.models.resnet_vae import BidirectionalResNetVAE, ModelError, ReparameterizedConv2D, make_coder


def test_reparameterized_conv_same_padding_and_data_init():
    torch.manual_seed(0)
    x = torch.randn(2, 3, 32, 32)
    conv = ReparameterizedConv2D(3, 8, (5, 5), (2, 2))
    y0 = conv(x)                                    # first call: "batch norm" initialisation (init_scale 0.1)
    assert y0.shape == (2, 8, 16, 16)
    assert torch.allclose(y0.mean(dim=(0, 2, 3)), torch.zeros(8), atol=1e-5)
    assert torch.allclose(y0.std(dim=(0, 2, 3), unbiased=False), torch.full((8,), 0.1), atol=1e-4)
    up = ReparameterizedConv2D(8, 3, (5, 5), (2, 2), transpose=True)
    assert up(y0).shape == (2, 3, 32, 32)
    same = ReparameterizedConv2D(8, 8, (3, 3))
    assert same(y0).shape == y0.shape


def test_signal_conv_sizes_and_gdn():
    x = torch.randn(1, 3, 64, 128)
    down = SignalConv2D(3, 6, (5, 5), corr=True, strides_down=2)
    y = down(x)
    assert y.shape == (1, 6, 32, 64)
    up = SignalConv2D(6, 3, (5, 5), strides_up=2)
    assert up(y).shape == (1, 3, 64, 128)
    assert SignalConv2D(6, 6, (3, 3), strides_up=1)(y).shape == y.shape
    g, ig = GDN(6, False), GDN(6, True)
    z = g(y)
    # at initialisation beta = 1, gamma = 0.1 I:  y / sqrt(1 + 0.1 y^2)
    assert torch.allclose(z, y * torch.rsqrt(1 + 0.1 * y * y), atol=1e-5)
    assert torch.allclose(ig(y), y * torch.sqrt(1 + 0.1 * y * y), atol=1e-5)


def test_rvae_latent_shapes_cpu():
    torch.manual_seed(0)
    model = BidirectionalResNetVAE(num_res_blocks=3, sampler="beam_search",
                                   sampler_args={"n_beams": 20, "extra_samples": 1.2}, coder_args={"block_size": 1000},
                                   deterministic_filters=16, stochastic_filters=32, kl_per_partition=3.)
    img = torch.rand(1, 32, 32, 3) - 0.5
    rec = model(img)
    assert rec.shape == (1, 32, 32, 3)
    for blk in model.residual_blocks:
        assert tuple(blk.posterior.loc.shape) == (1, 16, 16, 32) and tuple(blk.prior.scale.shape) == (1, 16, 16, 32)
        assert blk.coder.n_samples == 36 and blk.coder.n_beams == 20 and blk.coder.block_size == 1000
    assert float(model.kl_divergence()) > 0.
    with pytest.raises(ModelError):
        make_coder("nope", {}, {}, 3., "x")


def test_lossy_latent_shapes_cpu():
    torch.manual_seed(0)
    model = Large2LevelVAE(level_1_filters=12, level_2_filters=8)
    img = torch.rand(1, 128, 192, 3)
    rec = model(img)
    assert rec.shape == (1, 128, 192, 3)
    assert tuple(model.level_2_posterior.loc.shape) == (1, 2, 3, 8) == tuple(model.level_2_prior.loc.shape)
    assert tuple(model.level_1_posterior.loc.shape) == (1, 8, 12, 12) == tuple(model.level_1_prior.scale.shape)
    assert all(float(k) > 0. for k in model.kl_divergence())
