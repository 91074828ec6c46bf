"""Enums of the reference's public surface, with identical numeric values.

CompressionLevel.cs:9-80, CompressionStrategy.cs:9-37, FlushMode.cs:9-35, CompressionState.cs:9-55.
"""
import enum


class CompressionLevel(enum.IntEnum):
    DefaultCompression = -1
    Level0 = 0
    NoCompression = 0
    Level1 = 1
    BestSpeed = 1
    Level2 = 2
    Level3 = 3
    Level4 = 4
    Level5 = 5
    Level6 = 6
    Level7 = 7
    Level8 = 8
    Level9 = 9
    BestCompression = 9


class CompressionStrategy(enum.IntEnum):
    DefaultStrategy = 0
    Filtered = 1
    HuffmanOnly = 2
    Rle = 3
    Fixed = 4


class FlushMode(enum.IntEnum):
    NoFlush = 0
    PartialFlush = 1
    SyncFlush = 2
    FullFlush = 3
    Finish = 4


class CompressionState(enum.IntEnum):
    ZVERSIONERROR = -6
    ZBUFERROR = -5
    ZMEMERROR = -4
    ZDATAERROR = -3
    ZSTREAMERROR = -2
    ZERRNO = -1
    ZOK = 0
    ZSTREAMEND = 1
    ZNEEDDICT = 2


class HashVariant(enum.IntEnum):
    """Which UpdateHash the reference build would run (Deflate.Intrinsics.cs:295-307)."""
    Crc32C = 0      # >= netcoreapp3.1 on x86 with SSE4.2: the build the published numbers come from
    Multiplicative = 1
