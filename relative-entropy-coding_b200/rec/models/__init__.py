"""rec.models -- only the coder-facing glue of the reference's models: the NN layers are out of
scope, the compress/decompress loops around `coder.encode` / `coder.decode` / the `.rec` container are here."""
from .latent_hierarchy import LatentHierarchy, SyntheticLadder

__all__ = ["LatentHierarchy", "SyntheticLadder"]
