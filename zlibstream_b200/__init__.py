"""zlibstream_b200 -- B200-native segmented DEFLATE / INFLATE / Adler-32 behind the SixLabors.ZlibStream API.

The compute path is libzsb200.so (hand-written CUDA for sm_100a, C ABI in include/zsb200.h).  This package is
the host-side mirror of the reference's public surface; importing it does not need a GPU, using it does.
"""
from .enums import CompressionLevel, CompressionState, CompressionStrategy, FlushMode, HashVariant
from .options import ZlibOptions, ZlibStreamException

__all__ = ["CompressionLevel", "CompressionState", "CompressionStrategy", "FlushMode", "HashVariant", "ZlibOptions",
           "ZlibStreamException", "ZlibOutputStream", "ZlibInputStream", "Adler32", "Engine", "adler32_combine"]


def __getattr__(name):
    # the classes that need libzsb200.so are imported lazily so that `import zlibstream_b200` works anywhere
    if name in ("ZlibOutputStream", "ZlibInputStream"):
        from . import streams
        return getattr(streams, name)
    if name == "Adler32":
        from .adler32 import Adler32
        return Adler32
    if name in ("Engine", "adler32_combine"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
