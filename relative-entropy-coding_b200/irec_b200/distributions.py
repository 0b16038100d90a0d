"""Light stand-in for tfd.Normal: the coders only ever read `.loc` and `.scale`
(rec/coding/coder.py:427-431, beam_search_coder.py:53-77, samplers.py:74-96)."""
import torch


class Normal:
    """Diagonal Gaussian holder.  `loc` and `scale` are torch tensors of identical shape
    (anything array-like or any `__dlpack__` producer is converted); no sampling / log_prob is implemented on purpose --
    that arithmetic lives in the CUDA kernels."""

    def __init__(self, loc, scale, device=None):
        def conv(x):
            if not isinstance(x, torch.Tensor):
                if hasattr(x, "__dlpack__") and hasattr(x, "__dlpack_device__"):
                    x = torch.from_dlpack(x)       # any DLPack producer, zero-copy (data types)
                else:
                    if hasattr(x, "numpy"):
                        x = x.numpy()
                    x = torch.as_tensor(x)
            x = x.to(dtype=torch.float32)
            if device is not None:
                x = x.to(device)
            return x

        self.loc = conv(loc)
        self.scale = conv(scale)
        if self.loc.shape != self.scale.shape:
            self.loc, self.scale = torch.broadcast_tensors(self.loc, self.scale)
            self.loc, self.scale = self.loc.contiguous(), self.scale.contiguous()

    def __repr__(self):
        return f"Normal(shape={tuple(self.loc.shape)}, device={self.loc.device})"
