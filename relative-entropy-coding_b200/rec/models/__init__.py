"""rec.models -- the coder-facing side of the reference's models.

  * latent_hierarchy : the compress / decompress loops over an abstract ladder of priors/posteriors (synthetic stand-in
                       for the networks; what the benches use), for one image and for a batch of images (sub-batches
                       pipelined on CUDA streams: compress_batch / decompress_batch)
  * resnet_vae       : the lossless bidirectional ResNet VAE (rec/models/resnet_vae.py) in plain PyTorch, random-init
  * lossy            : the two-level lossy VAE (rec/models/lossy/large_2_level_vae.py) in plain PyTorch, random-init
The networks are callers of the hot path, not part of it: cuDNN convolutions, no custom kernels, no training loop."""
from .latent_hierarchy import BatchedSyntheticLadder, LatentHierarchy, SyntheticLadder
from .lossy import Large2LevelVAE
from .resnet_vae import BidirectionalResNetVAE

__all__ = ["LatentHierarchy", "SyntheticLadder", "BatchedSyntheticLadder", "BidirectionalResNetVAE", "Large2LevelVAE"]
