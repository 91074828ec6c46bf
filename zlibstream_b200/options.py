"""ZlibOptions (ZlibOptions.cs:14-24) and ZlibStreamException (ZlibStreamException.cs, ThrowHelper.cs:21-23)."""
from dataclasses import dataclass
from typing import Optional

from .enums import CompressionLevel, CompressionStrategy, FlushMode


@dataclass
class ZlibOptions:
    # CompressionLevel is None  =>  the stream inflates (ZlibStream.cs:20-28)
    CompressionLevel: Optional[CompressionLevel] = None
    CompressionStrategy: CompressionStrategy = CompressionStrategy.DefaultStrategy
    FlushMode: FlushMode = FlushMode.NoFlush


class ZlibStreamException(IOError):
    """The reference throws this IOException subclass for any engine state other than ZOK / ZSTREAMEND."""


def throw_compression_exception(compressing: bool, message: str):
    # ThrowHelper.cs:21-23: (compressing ? "de" : "in") + "flating: " + message
    raise ZlibStreamException(("de" if compressing else "in") + "flating: " + (message or ""))
