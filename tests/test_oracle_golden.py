"""Pins the CPU oracle to the reference: the 36 compressed sizes published in the reference's benchmarks.md
(tests/golden/sizes.json) and the properties its own tests check (round trip through an
independent inflater, ZlibStreamTests.Roundtrip.cs:44-63)."""
import hashlib
import json
import os
import zlib

import pytest

import corpus_data

HERE = os.path.dirname(os.path.abspath(__file__))
SIZES = json.load(open(os.path.join(HERE, "golden", "sizes.json")))
SEGMENTED = json.load(open(os.path.join(HERE, "golden", "segmented.json")))


@pytest.mark.parametrize("name", [k for k in SIZES if k != "sparse_rgba_3500"])
@pytest.mark.parametrize("level", [1, 3, 6])
def test_published_corpus_sizes(oracle, corpus, name, level):
    data = corpus[name]
    assert len(data) == SIZES[name]["size"]
    out, adler = oracle.deflate_oneshot(data, level)
    assert len(out) == SIZES[name][f"L{level}"]           # benchmarks.md `Bytes` column
    assert zlib.decompress(out) == data                    # independent inflater (SharpZipLib there, zlib here)
    assert adler == zlib.adler32(data)


@pytest.mark.parametrize("level", [1, 3, 6])
def test_published_sparse_image_sizes(oracle, level):
    data = corpus_data.sparse_image()                      # DeflateSparseBenchmark.cs:83-99
    assert len(data) == SIZES["sparse_rgba_3500"]["size"]
    out, _ = oracle.deflate_oneshot(data, level)
    assert len(out) == SIZES["sparse_rgba_3500"][f"L{level}"]
    assert zlib.decompress(out) == data


def test_multiplicative_hash_is_a_different_dialect(oracle, corpus):
    # Deflate.Intrinsics.cs:295-307: the non-SSE4.2 build hashes differently and compresses to different bytes
    a, _ = oracle.deflate_oneshot(corpus["alice29.txt"], 6, 0, 0)
    b, _ = oracle.deflate_oneshot(corpus["alice29.txt"], 6, 0, 1)
    assert len(a) == 55818 and a != b and zlib.decompress(b) == corpus["alice29.txt"]


def test_write_boundaries_change_the_stream(oracle, corpus):
    # the refill pre-insert makes the output depend on Write() boundaries even with NoFlush
    d = corpus["alice29.txt"]
    lib = oracle.lib
    sizes = {}
    for chunk in (1000, 4096, len(d)):
        h = lib.zso_deflater_new(6, 0, 0, 0)
        for i in range(0, len(d), chunk):
            assert lib.zso_deflater_write(h, d[i:i + chunk], len(d[i:i + chunk])) == 0
        assert lib.zso_deflater_finish(h) == 0
        sizes[chunk] = lib.zso_deflater_total_out(h)
        lib.zso_deflater_free(h)
    assert sizes == {1000: 55822, 4096: 55817, len(d): 55818}


@pytest.mark.parametrize("level,first2", [(0, "78da"), (1, "7801"), (2, "7801"), (3, "785e"), (4, "785e"), (5, "789c"),
                                          (6, "789c"), (7, "78da"), (8, "78da"), (9, "78da"), (-1, "789c")])
def test_header_flevel_is_old_zlib(oracle, level, first2):
    out, _ = oracle.deflate_oneshot(b"abc", level)        # Deflate.cs:466-480 (Q11)
    assert out[:2].hex() == first2


def test_empty_and_tiny_streams(oracle):
    out, adler = oracle.deflate_oneshot(b"", 6)
    assert out.hex() == "789c030000000001" and adler == 1
    for n in (1, 2, 3, 4, 261, 262, 263):
        d = bytes(range(256)) * 2
        d = d[:n]
        for level in (1, 6, 9):
            o, _ = oracle.deflate_oneshot(d, level)
            assert zlib.decompress(o) == d


@pytest.mark.parametrize("key", sorted(SEGMENTED))
def test_segmented_streams(oracle, corpus, key):
    """Full-flush-segmented streams (what the device path must reproduce): stable bytes, valid zlib, layout of
    SURVEY appendix B (every segment ends with the marker, the stream ends with 03 00 + Adler-32)."""
    _, seg, lvl, strat = key.split("/")
    ss, level, strategy = int(seg[3:]), int(lvl[1:]), int(strat[1:])
    data = b"".join(corpus[k] for k in sorted(corpus))
    out, ends, adler = oracle.deflate_segments(data, ss, level, strategy)
    g = SEGMENTED[key]
    assert len(out) == g["len"] and hashlib.sha256(out).hexdigest() == g["sha256"] and adler == g["adler"]
    assert hashlib.sha256(",".join(map(str, ends)).encode()).hexdigest() == g["ends_sha256"]
    assert zlib.decompress(out) == data
    assert all(out[e - 4:e] == b"\x00\x00\xff\xff" for e in ends)
    assert out[ends[-1]:ends[-1] + 2] == b"\x03\x00" and out[-4:] == adler.to_bytes(4, "big")
    if strategy != 3:
        # non-Rle segments decode on their own; Rle may reach one byte back (Q14)
        prev = 2
        for k, e in enumerate(ends[:8]):
            d = zlib.decompressobj(-15)
            assert d.decompress(out[prev:e]) == data[k * ss:(k + 1) * ss]
            prev = e


def test_first_segment_differs_from_later_ones(oracle, corpus):
    # identical content compresses differently as stream segment 0 and as segment k >= 1
    seg = corpus["lcet10.txt"][:65536]
    out, ends, _ = oracle.deflate_segments(seg * 3, 65536, 1)
    s0, s1, s2 = out[2:ends[0]], out[ends[0]:ends[1]], out[ends[1]:ends[2]]
    assert s1 == s2            # Q4: segments k >= 1 do not depend on what came before
    assert s0 != s1


def test_level0_and_strategies_round_trip(oracle, corpus):
    d = corpus["fields.c"]
    sizes = [len(oracle.deflate_oneshot(d, lvl)[0]) for lvl in range(10)]
    assert sizes == [11161, 3766, 3674, 3570, 3371, 3295, 3280, 3277, 3277, 3277]
    a = corpus["alice29.txt"]
    for (lvl, st, want) in [(6, 1, 55812), (6, 2, 87918), (6, 3, 87611), (1, 4, 85028), (6, 4, 71716)]:
        out, _ = oracle.deflate_oneshot(a, lvl, st)
        assert len(out) == want and zlib.decompress(out) == a


@pytest.mark.parametrize("level,strategy,ss", [(1, 0, 65536), (6, 0, 65536), (6, 3, 65536), (6, 2, 65536), (9, 1, 131072),
                                               (3, 0, 262144)])
def test_parallel_oracle_equals_the_single_thread_call(oracle, corpus, level, strategy, ss):
    """oracle_lib.deflate_segments_parallel (used by the multi-wave GPU parity tests and bench.py's parity check) must
    be the same function as zso_deflate_segments: ranges compressed behind their preceding segment."""
    blob = b"".join(corpus[k] for k in sorted(corpus))
    data = (blob + corpus_data.png_like(400000) + bytes(70000) + blob[::-1])[:22 * ss + 4321]
    want = oracle.deflate_segments(data, ss, level, strategy)
    got = oracle.deflate_segments_parallel(data, ss, level, strategy, threads=5)
    assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2]


def test_png_reference_round_trip():
    """oracle/png_ref.py (the checker of the device PNG front end): filter -> frame with stock zlib -> decode == pixels."""
    import struct
    import sys
    import zlib
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    import png_ref
    for ct in (0, 2, 4, 6):
        w, h, ch = 37, 19, png_ref.CHANNELS[ct]
        px = np.random.RandomState(ct).randint(0, 256, size=(h, w, ch)).astype(np.uint8).tobytes()
        for f in range(6):
            filt = png_ref.filter_image(px, w, h, ct, f)

            def chunk(t, d):
                return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

            png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ct, 0, 0, 0)) + \
                chunk(b"IDAT", zlib.compress(filt)) + chunk(b"IEND", b"")
            assert png_ref.decode(png)["pixels"] == px
