"""rec.models.lossy -- the two-level lossy VAE of the reference, as the CALLER of the coder.

Reference: rec/models/lossy/large_2_level_vae.py (transforms :13-251, Large2LevelVAE :254-456), layers
rec/models/custom_modules/signal_convolution.py (SignalConv2D: "same" reflect-padded strided correlation / transposed
convolution) and rec/models/custom_modules/gdn.py (GDN / inverse GDN).

Scope: glue around the hot path.  `compress` runs the inference pass, then the generative pass in
which `sampler.encode(target, coder, seed=seed)` codes level 2 and then level 1 (:343-385,:406-419), and writes the
`.rec` file; `decompress` reads the file and replays `sampler.decode(prior, seed=seed, indices=...)` per level
(:421-456).  Layers are plain PyTorch modules with random-init weights (no training loop, no DFT kernel
parametrisation -- an optimisation-time reparametrisation that does not change the function class).  NCHW inside the
network, the reference's NHWC `[1, H, W, C]` at the coder boundary.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from irec_b200.distributions import Normal
from rec.io.utils import read_compressed_code, write_compressed_code


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


class SignalConv2D(nn.Module):
    """signal_convolution.py:17-245 restricted to what the models use: padding="reflect", either strides_down (corr=True)
    or strides_up (corr=False); output size = input / strides_down resp. input * strides_up."""

    def __init__(self, in_channels, filters, kernel, corr=False, strides_down=1, strides_up=1, use_bias=True):
        super().__init__()
        self.k = tuple(kernel)
        assert self.k[0] % 2 == 1 and self.k[1] % 2 == 1, "odd kernels only"
        self.corr = bool(corr)
        self.down, self.up = int(strides_down), int(strides_up)
        fan_in = in_channels * self.k[0] * self.k[1]
        w = torch.randn(filters, in_channels, *self.k) * float(np.sqrt(1.0 / fan_in))      # variance scaling
        self.weight = nn.Parameter(w)
        self.bias = nn.Parameter(torch.zeros(filters)) if use_bias else None

    def forward(self, x):
        kh, kw = self.k
        if self.up == 1:
            # correlation with "same" reflect padding (k//2 before, (k-1)//2 after), then stride
            x = F.pad(x, (kw // 2, (kw - 1) // 2, kh // 2, (kh - 1) // 2), mode="reflect")
            w = self.weight if self.corr else torch.flip(self.weight, dims=(2, 3))
            return F.conv2d(x, w, self.bias, stride=self.down)
        # upsampling: zero-stuffed transposed convolution cropped to input * stride (signal_convolution.py:148-186)
        w = self.weight.permute(1, 0, 2, 3)
        ph, pw = (kh - 1) // 2, (kw - 1) // 2                 # odd kernels: (in - 1) * s - 2p + k + (s - 1) = in * s
        out = F.conv_transpose2d(x, w, self.bias, stride=self.up, padding=(ph, pw), output_padding=self.up - 1)
        return out


class GDN(nn.Module):
    """gdn.py:24-117: y = x * (beta + gamma * x^2)^(-1/2) (inverse: ^(+1/2)), non-negative reparametrisation"""

    def __init__(self, channels, inverse, gamma_init=0.1, beta_minimum=1e-6, gamma_minimum=0., reparam_offset=2. ** -18):
        super().__init__()
        self.inverse = bool(inverse)
        self.pedestal = reparam_offset ** 2
        self.beta_bound = (beta_minimum + self.pedestal) ** 0.5
        self.gamma_bound = (gamma_minimum + self.pedestal) ** 0.5
        self._beta = nn.Parameter(torch.sqrt(torch.ones(channels) + self.pedestal))
        self._gamma = nn.Parameter(torch.sqrt(gamma_init * torch.eye(channels) + self.pedestal))

    def forward(self, x):
        beta = torch.clamp_min(self._beta, self.beta_bound) ** 2 - self.pedestal
        gamma = torch.clamp_min(self._gamma, self.gamma_bound) ** 2 - self.pedestal
        norm = F.conv2d(x * x, gamma.t().reshape(gamma.shape[1], gamma.shape[0], 1, 1), beta)
        return x * (torch.sqrt(norm) if self.inverse else torch.rsqrt(norm))


class AnalysisTransform(nn.Module):            # :13-80
    def __init__(self, num_filters):
        super().__init__()
        n = num_filters
        self.layers = nn.Sequential(
            SignalConv2D(3, n, (5, 5), corr=True, strides_down=2), GDN(n, False),
            SignalConv2D(n, n, (5, 5), corr=True, strides_down=2), GDN(n, False),
            SignalConv2D(n, n, (5, 5), corr=True, strides_down=2), GDN(n, False))
        self.loc_head = SignalConv2D(n, n, (5, 5), corr=True, strides_down=2)
        self.log_scale_head = SignalConv2D(n, n, (5, 5), corr=True, strides_down=2)

    def forward(self, x):
        x = self.layers(x)
        return self.loc_head(x), self.log_scale_head(x)


class SynthesisTransform(nn.Module):           # :83-134
    def __init__(self, num_filters):
        super().__init__()
        n = num_filters
        self.layers = nn.Sequential(
            SignalConv2D(n, n, (5, 5), strides_up=2), GDN(n, True),
            SignalConv2D(n, n, (5, 5), strides_up=2), GDN(n, True),
            SignalConv2D(n, n, (5, 5), strides_up=2), GDN(n, True),
            SignalConv2D(n, 3, (5, 5), strides_up=2))

    def forward(self, x):
        return self.layers(x)


class HyperAnalysisTransform(nn.Module):       # :137-189
    def __init__(self, in_filters, num_filters):
        super().__init__()
        n = num_filters
        self.conv0 = SignalConv2D(in_filters, n, (3, 3), corr=True, strides_down=1)
        self.conv1 = SignalConv2D(n, n, (5, 5), corr=True, strides_down=2)
        self.loc_head = SignalConv2D(n, n, (5, 5), corr=True, strides_down=2, use_bias=False)
        self.log_scale_head = SignalConv2D(n, n, (5, 5), corr=True, strides_down=2, use_bias=False)

    def forward(self, x):
        x = F.relu(self.conv1(F.relu(self.conv0(x))))
        return self.loc_head(x), self.log_scale_head(x)


class HyperSynthesisTransform(nn.Module):      # :192-251
    def __init__(self, num_filters, num_output_filters):
        super().__init__()
        n = num_filters
        self.conv0 = SignalConv2D(n, n, (5, 5), strides_up=2)
        self.conv1 = SignalConv2D(n, n, (5, 5), strides_up=2)
        self.loc_head = SignalConv2D(n, num_output_filters, (3, 3), strides_up=1)
        self.log_scale_head = SignalConv2D(n, num_output_filters, (3, 3), strides_up=1)

    def forward(self, x):
        x = F.relu(self.conv1(F.relu(self.conv0(x))))
        return self.loc_head(x), self.log_scale_head(x)


class Large2LevelVAE(nn.Module):
    """large_2_level_vae.py:254-456.  `sampler` in compress/decompress is the coder object of the examples
    (compress_with_lossy_model.py:110-124), e.g. a BeamSearchCoder with block_size."""

    def __init__(self, level_1_filters=196, level_2_filters=128, name="large_2_level_vae"):
        super().__init__()
        l1, l2 = level_1_filters, level_2_filters
        self.level_1_filters, self.level_2_filters = l1, l2
        self._prior_base = nn.Parameter(torch.zeros(1, l2, 1, 1))
        self._prior_conv = SignalConv2D(l2, l2, (3, 3), corr=True)
        self._prior_loc_head = SignalConv2D(l2, l2, (3, 3), corr=True)
        self._prior_log_scale_head = SignalConv2D(l2, l2, (3, 3), corr=True)
        self._level_1_posterior_loc_combiner = nn.Conv2d(2 * l1, l1, 1)
        self._level_1_posterior_log_scale_combiner = nn.Conv2d(2 * l1, l1, 1)
        self.analysis_transform = AnalysisTransform(l1)
        self.synthesis_transform = SynthesisTransform(l1)
        self.hyper_analysis_transform = HyperAnalysisTransform(l1, l2)
        self.hyper_synthesis_transform = HyperSynthesisTransform(l2, l1)
        self.level_1_prior = self.level_1_posterior = self.level_2_prior = self.level_2_posterior = None

    def prior_base(self, batch_size, height, width):
        return self._prior_base.expand(batch_size, -1, height // 64, width // 64).contiguous()

    def _level_2_prior(self, batch_size, height, width):                      # :335-341 / :432-439
        t = F.elu(self._prior_conv(self.prior_base(batch_size, height, width)))
        return Normal(_nhwc(self._prior_loc_head(t)), _nhwc(F.softplus(self._prior_log_scale_head(t)) + 1e-7))

    def _level_1_prior(self, level_2_latent_nhwc):                            # :352-353 / :444-448
        loc, log_scale = self.hyper_synthesis_transform(_nchw(level_2_latent_nhwc))
        return loc, log_scale, Normal(_nhwc(loc), _nhwc(F.softplus(log_scale) + 1e-7))

    def kl_divergence(self):
        def kl(q, p):
            dl = torch.log(q.scale) - torch.log(p.scale)
            return (0.5 * ((q.loc - p.loc) / p.scale) ** 2 + 0.5 * torch.expm1(2. * dl) - dl).sum()
        return [kl(self.level_1_posterior, self.level_1_prior), kl(self.level_2_posterior, self.level_2_prior)]

    @torch.no_grad()
    def forward(self, image, sampling_fn=None):
        """image: [N, H, W, 3]; sampling_fn(target=, coder=) -> (indices, sample) or None to draw from the posteriors
        (:305-404)"""
        x = _nchw(image)
        n, _, h, w = x.shape
        l1_post_loc, l1_post_log_scale = self.analysis_transform(x)
        l2_post_loc, l2_post_log_scale = self.hyper_analysis_transform(l1_post_loc)
        self.level_2_posterior = Normal(_nhwc(l2_post_loc), _nhwc(F.softplus(l2_post_log_scale) + 1e-7))
        self.level_2_prior = self._level_2_prior(n, h, w)
        if sampling_fn is None:
            q = self.level_2_posterior
            z2 = q.loc + q.scale * torch.randn_like(q.loc)
        else:
            level_2_indices, z2 = sampling_fn(target=self.level_2_posterior, coder=self.level_2_prior)
        l1_prior_loc, l1_prior_log_scale, self.level_1_prior = self._level_1_prior(z2)
        loc = F.elu(torch.cat([l1_post_loc, l1_prior_loc], dim=1))
        log_scale = F.elu(torch.cat([l1_post_log_scale, l1_prior_log_scale], dim=1))
        loc = self._level_1_posterior_loc_combiner(loc)
        scale = F.softplus(self._level_1_posterior_log_scale_combiner(log_scale)) + 1e-7
        self.level_1_posterior = Normal(_nhwc(loc), _nhwc(scale))
        if sampling_fn is None:
            q = self.level_1_posterior
            z1 = q.loc + q.scale * torch.randn_like(q.loc)
        else:
            level_1_indices, z1 = sampling_fn(target=self.level_1_posterior, coder=self.level_1_prior)
        rec = _nhwc(self.synthesis_transform(_nchw(z1)))
        if sampling_fn is None:
            return rec
        return [level_2_indices, level_1_indices], rec

    # -- :406-419 -------------------------------------------------------------------------------------------------------
    def compress(self, file_path, image, seed, sampler, block_size, max_index):
        """image: [H, W, 3]"""
        sampling_fn = lambda target, coder: sampler.encode(target, coder, seed=seed)          # noqa: E731
        block_indices, reconstruction = self(image[None, ...], sampling_fn=sampling_fn)
        block_indices = [[[int(i) for i in blk] for blk in level] for level in block_indices]
        write_compressed_code(file_path=file_path, seed=seed, image_shape=tuple(image.shape), block_size=block_size,
                              block_indices=block_indices, max_index=max_index)
        return reconstruction

    # -- :421-456 -------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def decompress(self, file_path, sampler):
        seed, image_shape, block_size, block_indices = read_compressed_code(file_path=file_path)
        height, width = image_shape[0], image_shape[1]
        self.level_2_prior = self._level_2_prior(1, height, width)
        z2 = sampler.decode(self.level_2_prior, seed=seed, indices=[list(b) for b in block_indices[0]])
        _, _, self.level_1_prior = self._level_1_prior(z2)
        z1 = sampler.decode(self.level_1_prior, seed=seed, indices=[list(b) for b in block_indices[1]])
        return _nhwc(self.synthesis_transform(_nchw(z1)))
