"""rec.io -- wire format of the iREC index stream (reference: rec/io/, whose __init__.py is empty; the two container
functions are re-exported here because examples/lossless/compression_performance.py:14 imports them from `rec.io`)."""
from .utils import read_compressed_code, write_compressed_code
from .entropy_coding import ArithmeticCoder

__all__ = ["ArithmeticCoder", "write_compressed_code", "read_compressed_code"]
