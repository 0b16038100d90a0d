"""irec_b200: B200-native engine behind the reference's rec.coding API (see ../rec/coding)."""
from .distributions import Normal
from .engine import CodingError
from . import native

__all__ = ["Normal", "CodingError", "native"]
