"""the reference's own call pattern -- one Write with FlushMode.NoFlush, then Dispose -- on the device.

The 36 compressed sizes of the reference's benchmarks.md (tests/golden/sizes.json: Canterbury files + the sparse image, levels
1 / 3 / 6) are produced by exactly this call, so here they pin the CUDA path itself, not only the oracle.  Bytes are compared
with the oracle's one-shot stream; every stream is inflated by the device (one-piece streams end with a BFINAL block) and by
stock zlib."""
import json
import os
import zlib

import pytest

import corpus_data

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SIZES = json.load(open(os.path.join(HERE, "golden", "sizes.json")))


@pytest.mark.parametrize("name", sorted(corpus_data.load_corpus()))
@pytest.mark.parametrize("level", [1, 3, 6])
def test_published_sizes_on_the_device(engine, oracle, corpus, name, level):
    data = corpus[name]
    got, adler = engine.deflate_stream(data, level)
    assert len(got) == SIZES[name][f"L{level}"]          # the reference's own number (benchmarks.md)
    want, want_adler = oracle.deflate_oneshot(data, level)
    assert got == want and adler == want_adler == zlib.adler32(data)
    assert zlib.decompress(got) == data
    dec, ad = engine.inflate_stream(got, len(data))
    assert dec == data and ad == adler


@pytest.mark.parametrize("level,strategy", [(9, 0), (4, 0), (6, 1), (6, 2), (6, 3), (1, 4), (2, 0)])
def test_one_piece_stream_other_settings(engine, oracle, corpus, level, strategy):
    for name in ("alice29.txt", "ptt5", "kennedy.xls"):
        data = corpus[name]
        got, adler = engine.deflate_stream(data, level, strategy)
        want, _ = oracle.deflate_oneshot(data, level, strategy)
        assert got == want, (name, level, strategy)
        dec, ad = engine.inflate_stream(got, len(data))
        assert dec == data and ad == adler


@pytest.mark.parametrize("n", [0, 1, 2, 3, 258, 65535, 65536, 65537, 98304, 163840, 200001])
def test_one_piece_stream_sizes(engine, oracle, n):
    data = corpus_data.text_like(n, seed=n + 1) if n else b""
    got, adler = engine.deflate_stream(data, 6)
    want, _ = oracle.deflate_oneshot(data, 6)
    assert got == want
    assert zlib.decompress(got) == data
    dec, ad = engine.inflate_stream(got, len(data))
    assert dec == data and ad == (zlib.adler32(data) & 0xFFFFFFFF)


def test_sparse_image_published_sizes(engine, oracle):
    data = corpus_data.sparse_image()
    for level, want_len in ((1, 825050), (6, 659280)):     # benchmarks.md:18,28
        got, _ = engine.deflate_stream(data, level)
        assert len(got) == want_len
        assert zlib.decompress(got) == data


@pytest.mark.parametrize("zlevel", [1, 6, 9])
def test_device_inflates_a_third_party_one_piece_stream(engine, corpus, zlevel):
    """A stream no part of this repo wrote: stock zlib's (stored, static and dynamic blocks, its own block boundaries)."""
    data = corpus["lcet10.txt"] + bytes(70000) + corpus["kennedy.xls"][:300000] + os.urandom(50000)
    comp = zlib.compress(data, zlevel)
    dec, ad = engine.inflate_stream(comp, len(data))
    assert dec == data and ad == (zlib.adler32(data) & 0xFFFFFFFF)
    bad = bytearray(comp)
    bad[-1] ^= 1
    with pytest.raises(Exception, match="incorrect data check"):
        engine.inflate_stream(bytes(bad), len(data))
