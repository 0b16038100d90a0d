"""rec.models.resnet_vae -- the lossless bidirectional ResNet VAE of the reference, as the CALLER of the coder.

Reference: rec/models/resnet_vae.py (BidirectionalResidualBlock :22-495, BidirectionalResNetVAE :498-860) with the
weight-normalised convolutions of rec/models/custom_modules/reparameterized_convolutions.py:55-290.

Scope: this is glue around the hot path -- the coder construction from strings (:121-141), the
inference pass that fixes the posteriors' first halves, and the sequential generative pass in which every block calls
`coder.encode(posterior, prior, seed=...)` / `coder.decode(prior, seed=..., indices=...)` (:470,:476,:803-860).  The
layers are plain PyTorch modules (cuDNN convolutions; no custom kernels, no training loop, no IAF posterior); weights
are random-init exactly as the reference builds them (N(0, 0.05) direction tensors, data-dependent scale/bias
initialisation on the first forward pass, :236-258 of the convolution module).  Tensors are NCHW inside the network
and handed to the coder as the reference's NHWC `[1, H, W, C]` so that the flattening order -- hence which dims share
a coder-block after `Coder.split` -- is the reference's.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from irec_b200.distributions import Normal
from rec.coding import BeamSearchCoder, GaussianCoder
from rec.coding.samplers import ImportanceSampler, RejectionSampler


class ModelError(Exception):
    pass


class ReparameterizedConv2D(nn.Module):
    """weight-normalised convolution: kernel = l2_normalize(V over (h, w, in)) * exp(log_scale)
    (reparameterized_convolutions.py:115-122); the first call initialises log_scale and bias from the batch moments of
    its own output (:236-258, "batch norm" initialisation with init_scale 0.1; the stored value is log(scale) / 3 as in
    the reference, :252)."""

    def __init__(self, in_channels, filters, kernel_size, strides=(1, 1), transpose=False, use_bias=True):
        super().__init__()
        kh, kw = kernel_size
        self.strides = tuple(strides)
        self.transpose = transpose
        self.kernel_size = (kh, kw)
        shape = (in_channels, filters, kh, kw) if transpose else (filters, in_channels, kh, kw)
        self.kernel_weights = nn.Parameter(0.05 * torch.randn(shape))
        self.kernel_log_scale = nn.Parameter(torch.zeros(filters))
        self.bias = nn.Parameter(torch.zeros(filters)) if use_bias else None
        self.register_buffer("_initialized", torch.tensor(False))

    def _kernel(self, initializing=False):
        out_axis = 1 if self.transpose else 0
        dims = tuple(d for d in range(4) if d != out_axis)
        v = self.kernel_weights / self.kernel_weights.pow(2).sum(dim=dims, keepdim=True).clamp_min(1e-12).sqrt()
        if not initializing:
            shape = [1, 1, 1, 1]
            shape[out_axis] = -1
            v = v * torch.exp(self.kernel_log_scale).reshape(shape)
        return v

    def _conv(self, x, kernel):
        kh, kw = self.kernel_size
        sh, sw = self.strides
        if self.transpose:
            # TF "same" transposed convolution: output = input * stride
            ph, pw = (kh - sh + 1) // 2, (kw - sw + 1) // 2
            oph, opw = 2 * ph + sh - kh, 2 * pw + sw - kw
            return F.conv_transpose2d(x, kernel, stride=(sh, sw), padding=(ph, pw), output_padding=(oph, opw))
        # TF "same": total padding = max(k - s, 0) when the size is a multiple of the stride, extra pixel at the end
        H, W = x.shape[-2:]
        th = max((-(-H // sh) - 1) * sh + kh - H, 0)
        tw = max((-(-W // sw) - 1) * sw + kw - W, 0)
        x = F.pad(x, (tw // 2, tw - tw // 2, th // 2, th - th // 2))
        return F.conv2d(x, kernel, stride=(sh, sw))

    def forward(self, x, init_scale=0.1):
        if not bool(self._initialized):
            with torch.no_grad():
                out = self._conv(x, self._kernel(initializing=True))
                mean = out.mean(dim=(0, 2, 3), keepdim=True)
                var = out.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
                scale_init = init_scale / torch.sqrt(var + 1e-10)
                self.kernel_log_scale.copy_((torch.log(scale_init) / 3.0).reshape(-1))
                if self.bias is not None:
                    self.bias.copy_((-mean * scale_init).reshape(-1))
                self._initialized.fill_(True)
                return (out - mean) * scale_init
        out = self._conv(x, self._kernel())
        if self.bias is not None:
            out = out + self.bias.reshape(1, -1, 1, 1)
        return out


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def make_coder(sampler, sampler_args, coder_args, kl_per_partition, name):
    """the coder construction of BidirectionalResidualBlock.__init__ (resnet_vae.py:121-141)"""
    if sampler == "rejection":
        return GaussianCoder(sampler=RejectionSampler(**sampler_args), kl_per_partition=kl_per_partition,
                             name=f"encoder_for_{name}", **coder_args)
    if sampler == "importance":
        return GaussianCoder(sampler=ImportanceSampler(**sampler_args), kl_per_partition=kl_per_partition,
                             name=f"encoder_for_{name}", **coder_args)
    if sampler == "beam_search":
        return BeamSearchCoder(kl_per_partition=kl_per_partition, n_beams=sampler_args["n_beams"],
                               extra_samples=sampler_args["extra_samples"], name=f"encoder_for_{name}", **coder_args)
    raise ModelError(f"Sampler must be one of ['rejection', 'importance', 'beam_search'],but got {sampler}!")


class BidirectionalResidualBlock(nn.Module):
    """resnet_vae.py:22-495 (Gaussian latents, no IAF)"""

    def __init__(self, stochastic_filters, deterministic_filters, sampler, sampler_args=None, coder_args=None,
                 kernel_size=(3, 3), is_last=False, kl_per_partition=8., name="bidirectional_resnet_block"):
        super().__init__()
        self.name = name
        self.is_last = is_last
        sf, df, k = stochastic_filters, deterministic_filters, kernel_size
        conv = lambda cin, cout: ReparameterizedConv2D(cin, cout, k)      # noqa: E731
        if not is_last:
            self.infer_conv1 = conv(df, df)
            self.infer_conv2 = conv(df, df)
        self.infer_posterior_loc_head = conv(df, sf)
        self.infer_posterior_log_scale_head = conv(df, sf)
        self.gen_conv1 = conv(df, df)
        self.gen_conv2 = conv(df + sf, df)
        self.prior_loc_head = conv(df, sf)
        self.prior_log_scale_head = conv(df, sf)
        self.gen_posterior_loc_head = conv(df, sf)
        self.gen_posterior_log_scale_head = conv(df, sf)
        self.coder = make_coder(sampler, dict(sampler_args or {}), dict(coder_args or {}), kl_per_partition, name)
        self.infer_posterior_loc = 0.
        self.infer_posterior_log_scale = 0.
        self.posterior = None
        self.prior = None
        self.register_buffer("_initialized", torch.tensor(False))

    def forward(self, tensor, inference_pass=True, encoder_args=None, decoder_args=None):
        inp = tensor
        tensor = F.elu(tensor)
        indices = None
        if inference_pass:                                                   # :387-401
            self.infer_posterior_loc = self.infer_posterior_loc_head(tensor)
            self.infer_posterior_log_scale = self.infer_posterior_log_scale_head(tensor)
            if not self.is_last:
                tensor = self.infer_conv2(F.elu(self.infer_conv1(tensor)))
        else:                                                                # :406-487
            prior_loc = self.prior_loc_head(tensor)
            prior_scale = torch.exp(self.prior_log_scale_head(tensor))
            self.prior = Normal(_nhwc(prior_loc), _nhwc(prior_scale))
            if decoder_args is None:
                post_loc = self.infer_posterior_loc + self.gen_posterior_loc_head(tensor)
                post_scale = torch.exp(self.infer_posterior_log_scale + self.gen_posterior_log_scale_head(tensor))
                self.posterior = Normal(_nhwc(post_loc), _nhwc(post_scale))
            if encoder_args is None and decoder_args is None:                # training-style pass (:418-454)
                if bool(self._initialized):
                    latent = post_loc + post_scale * torch.randn_like(post_loc)
                else:
                    latent = prior_loc + prior_scale * torch.randn_like(prior_loc)
                    self._initialized.fill_(True)
            elif encoder_args is not None:                                   # compression (:459-470)
                args = dict(encoder_args)
                max_aux = args.pop("max_aux", None)
                if post_loc.shape[0] > 1:
                    # a batch of images (extension, BASELINE.json configs[3]): every image coded independently with the same
                    # seed, ONE launch for the level; `indices` is a callable -> [image][coder_block] lists
                    if not hasattr(self.coder, "encode_batch"):
                        raise ModelError("compress_batch needs a coder with encode_batch (beam_search)")
                    indices, latent_nhwc = self.coder.encode_batch(self.posterior, self.prior, seed=args["seed"], lazy=True)
                elif max_aux is not None and hasattr(self.coder, "encode_lazy"):
                    # no host synchronisation: `indices` is a callable that reads the lists back later
                    indices, latent_nhwc = self.coder.encode_lazy(self.posterior, self.prior, seed=args["seed"], max_aux=max_aux)
                else:
                    indices, latent_nhwc = self.coder.encode(self.posterior, self.prior, **args)
                latent = _nchw(latent_nhwc)
            elif prior_loc.shape[0] > 1:                                     # decompression of a batch (extension)
                latent = _nchw(self.coder.decode_batch(self.prior, decoder_args["indices"], seed=decoder_args["seed"]))
            else:                                                            # decompression (:475-476)
                latent = _nchw(self.coder.decode(self.prior, **decoder_args))
            tensor = self.gen_conv1(tensor)
            tensor = torch.cat([tensor, latent], dim=1)
            tensor = self.gen_conv2(F.elu(tensor))
        tensor = inp + 0.1 * tensor
        if encoder_args is not None:
            return indices, tensor
        return tensor

    def kl_divergence(self):
        q, p = self.posterior, self.prior
        dl = torch.log(q.scale) - torch.log(p.scale)
        return (0.5 * ((q.loc - p.loc) / p.scale) ** 2 + 0.5 * torch.expm1(2. * dl) - dl).sum()


class BidirectionalResNetVAE(nn.Module):
    """resnet_vae.py:498-860: compress / decompress / get_codelength over `num_res_blocks` stochastic levels"""

    def __init__(self, num_res_blocks, sampler, sampler_args=None, coder_args=None, first_kernel_size=(5, 5),
                 first_strides=(2, 2), kernel_size=(3, 3), deterministic_filters=160, stochastic_filters=32,
                 kl_per_partition=8., name="resnet_vae"):
        super().__init__()
        self.num_res_blocks = num_res_blocks
        self.deterministic_filters = deterministic_filters
        self.stochastic_filters = stochastic_filters
        self.first_strides = tuple(first_strides)
        self.likelihood_log_scale = nn.Parameter(torch.zeros(()))
        self.first_infer_conv = ReparameterizedConv2D(3, deterministic_filters, first_kernel_size, first_strides)
        self.last_gen_conv = ReparameterizedConv2D(deterministic_filters, 3, first_kernel_size, first_strides, transpose=True)
        self.residual_blocks = nn.ModuleList([
            BidirectionalResidualBlock(stochastic_filters, deterministic_filters, sampler, sampler_args, coder_args,
                                       kernel_size=kernel_size, is_last=(i == 0), kl_per_partition=kl_per_partition,
                                       name=f"resnet_block_{i}") for i in range(num_res_blocks)])
        self._generative_base = nn.Parameter(torch.zeros(deterministic_filters))
        self.log_likelihood = -np.inf

    def generative_base(self, batch_size, height, width):
        sh, sw = self.first_strides
        return self._generative_base.reshape(1, -1, 1, 1).expand(batch_size, -1, height // sh, width // sw).contiguous()

    def _likelihood(self, reference, reconstruction, binsize=1. / 256.):     # discretized logistic, :624-636
        scale = torch.exp(self.likelihood_log_scale)
        x = (torch.floor(reference / binsize) * binsize - reconstruction) / scale
        ll = torch.sigmoid(x + binsize / scale) - torch.sigmoid(x)
        return torch.log(ll + 1e-7).sum(dim=(1, 2, 3))

    def _reconstruct(self, tensor):
        rec = self.last_gen_conv(F.elu(tensor))
        return torch.clamp(rec, -0.5 + 1. / 512., 0.5 - 1. / 512.)

    @torch.no_grad()
    def forward(self, image):
        """image: [N, H, W, 3] in [-0.5, 0.5] (reference :676-716); also performs the data-dependent initialisation"""
        x = _nchw(image)
        n, _, h, w = x.shape
        t = self.first_infer_conv(x)
        for blk in reversed(self.residual_blocks):
            t = blk(t, inference_pass=True)
        t = self.generative_base(n, h, w)
        for blk in self.residual_blocks:
            t = blk(t, inference_pass=False)
        rec = self._reconstruct(t)
        self.log_likelihood = self._likelihood(x, rec).mean()
        return _nhwc(rec) + 0.5

    def kl_divergence(self):
        return sum(blk.kl_divergence() for blk in self.residual_blocks)

    # -- :803-836 -------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def compress(self, image, seed, update_sampler=False, max_aux=None, _again_done=False):
        """max_aux (extension): promise that no coder-block needs more auxiliary variables than this; the blocks are
        then coded without a host synchronisation between them (BeamSearchCoder.encode_lazy) and the index lists are
        read back after the last block; CodingError if the promise is broken."""
        x = _nchw(image)
        n, _, h, w = x.shape
        t = self.first_infer_conv(x)
        for blk in reversed(self.residual_blocks):
            t = blk(t, inference_pass=True)
        t = self.generative_base(n, h, w)
        block_indices = []
        enc_args = {"seed": seed, "update_sampler": update_sampler}
        if max_aux is not None:
            enc_args["max_aux"] = max_aux
        for blk in self.residual_blocks:
            indices, t = blk(t, inference_pass=False, encoder_args=enc_args)
            block_indices.append(indices)
        pending = block_indices
        block_indices = [ind() if callable(ind) else ind for ind in pending]
        if any(getattr(ind, "retried", False) for ind in pending) and not _again_done:
            # a batched launch ran out of index rows and was repeated after later blocks had consumed its latent: the row
            # capacity hint has grown, one more pass runs without repeats (engine.PendingBeamResult)
            return self.compress(image, seed, update_sampler=update_sampler, max_aux=max_aux, _again_done=True)
        rec = self._reconstruct(t)
        self.log_likelihood = self._likelihood(x, rec).mean()
        return block_indices, _nhwc(rec)

    @torch.no_grad()
    def compress_batch(self, images, seed):
        """BASELINE.json configs[3]: a batch of images [N, H, W, 3] through the model, every image coded independently (what
        looping `compress` over them computes, up to the convolutions' batch-size dependent rounding): one coder launch per
        residual block for the whole batch, no host synchronisation between the blocks, index lists read back at the end.
        Returns (block_indices[image][res_block][coder_block], reconstructions [N, H, W, 3])."""
        per_level, rec = self.compress(images, seed)
        n = images.shape[0]
        if n == 1:
            return [per_level], rec
        return [[per_level[k][i] for k in range(len(per_level))] for i in range(n)], rec

    @torch.no_grad()
    def decompress_batch(self, compressed_codes, seed, height=32, width=32):
        """compressed_codes[image][res_block][coder_block] as returned by compress_batch -> images [N, H, W, 3] in [0, 1]"""
        n = len(compressed_codes)
        if n == 1:
            return self.decompress(compressed_codes[0], seed, height=height, width=width)
        t = self.generative_base(n, height, width)
        for k, blk in enumerate(self.residual_blocks):
            codes = [[list(b) for b in compressed_codes[i][k]] for i in range(n)]
            t = blk(t, inference_pass=False, decoder_args={"seed": seed, "indices": codes})
        return _nhwc(self._reconstruct(t)) + 0.5

    # -- :838-842 -------------------------------------------------------------------------------------------------------
    def get_codelength(self, compressed_codes):
        total = 0.
        for blk, code in zip(self.residual_blocks, compressed_codes):
            total += blk.coder.get_codelength(code)
        return total

    # -- :844-860 (the reference hard-codes a 32x32 image there; height/width are parameters here) -------------------------
    @torch.no_grad()
    def decompress(self, compressed_codes, seed, height=32, width=32, lossless=True):
        t = self.generative_base(1, height, width)
        for blk, code in zip(self.residual_blocks, compressed_codes):
            t = blk(t, inference_pass=False, decoder_args={"seed": seed, "indices": code})
        return _nhwc(self._reconstruct(t)) + 0.5
