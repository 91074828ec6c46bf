"""Oracle inflate / Adler-32 against system zlib (the independent oracle the reference's tests use SharpZipLib for)
and the reference's own Adler32Tests.cs cases."""
import random
import zlib

import numpy as np
import pytest


def _rand(n, seed=1):
    return np.random.RandomState(seed).bytes(n)


@pytest.mark.parametrize("n", [0, 8, 215, 1024, 1024 + 15, 2034, 4096, 5552, 5553, 100000])
def test_adler_matches_reference_lengths(oracle, n):
    d = _rand(n)                                           # Adler32Tests.cs:22-40
    assert oracle.adler32(d) == zlib.adler32(d)
    assert oracle.adler32(d, 0x00AB00CD) == zlib.adler32(d, 0x00AB00CD)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_adler_empty_returns_seed(oracle, seed):
    assert oracle.adler32(b"", seed) == seed               # Adler32Tests.cs:13-20


def test_adler_combine(oracle):
    rng = random.Random(3)
    for _ in range(200):
        a = _rand(rng.randrange(0, 70000), rng.randrange(1 << 30))
        b = _rand(rng.randrange(0, 70000), rng.randrange(1 << 30))
        assert oracle.adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(a + b)


def test_inflate_decodes_stock_zlib_streams(oracle, corpus):
    for name in ("alice29.txt", "kennedy.xls", "ptt5", "sum", "xargs.1"):
        d = corpus[name]
        for level in (0, 1, 6, 9):
            comp = zlib.compress(d, level)
            rc, out, adler, msg, used = oracle.inflate(comp, len(d))
            assert rc == 1 and out == d and adler == zlib.adler32(d) and used == len(comp), (name, level, msg)
    for strategy in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
        c = zlib.compressobj(6, zlib.DEFLATED, 15, 8, strategy)
        d = corpus["cp.html"]
        comp = c.compress(d) + c.flush()
        assert oracle.inflate(comp, len(d))[1] == d


def test_inflate_round_trips_the_oracle_deflate(oracle, corpus):
    d = b"".join(corpus[k] for k in sorted(corpus))
    for level, strat in [(1, 0), (6, 0), (6, 3), (1, 4)]:
        comp, _, adler = oracle.deflate_segments(d, 65536, level, strat)
        rc, out, ad, msg, _ = oracle.inflate(comp, len(d))
        assert rc == 1 and out == d and ad == adler, msg


def test_inflate_error_vocabulary(oracle):
    good = zlib.compress(b"hello hello hello hello", 6)
    cases = [
        (b"\x79" + good[1:], "unknown compression method"),        # Inflate.cs:134
        (b"\x88\x1c" + good[2:], "invalid window size"),             # Inflate.cs:142
        (good[:1] + b"\x00" + good[2:], "incorrect header check"),   # Inflate.cs:166
        (good[:-1] + bytes([good[-1] ^ 1]), "incorrect data check"), # Inflate.cs:339
        (b"\x78\x9c\x07", "invalid block type"),                     # InflateBlocks.cs:237
        (b"\x78\x9c\x01\x05\x00\x00\x00", "invalid stored block lengths"),  # InflateBlocks.cs:278
    ]
    for comp, want in cases:
        rc, _, _, msg, _ = oracle.inflate(comp, 1000)
        assert rc == -3 and msg == want, (comp.hex(), msg, want)
    rc, _, _, msg, _ = oracle.inflate(b"\x78\xbb" + good[2:], 1000)   # FDICT set (Inflate.cs:171-243)
    assert rc == 2 and msg == "need dictionary"
    rc, _, _, msg, _ = oracle.inflate(good[:-3], 1000)                # truncated
    assert rc == -5
