"""GPU inflate of full-flush-segmented streams, Adler-32, and the stream classes, all through the C ABI."""
import io
import zlib

import numpy as np
import pytest

import corpus_data

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["auto", "warp", "lanes", "lanes_small"], autouse=True)
def inflate_kernel(request, monkeypatch):
    """The library picks the warp-per-segment or the lane-per-segment inflate kernel from the number of segments; the
    parity tests run with the automatic choice and with each kernel forced (ZSB_INFLATE_KERNEL is read per call)."""
    if request.param == "auto":
        monkeypatch.delenv("ZSB_INFLATE_KERNEL", raising=False)
    else:
        monkeypatch.setenv("ZSB_INFLATE_KERNEL", request.param)
    return request.param


@pytest.mark.parametrize("n", [1, 8, 215, 1024, 1024 + 15, 2034, 4096, 5552, 5553, 65535, 65536, 65537, (1 << 22) + 13])
def test_adler32_matches_zlib(engine, n):
    d = np.random.RandomState(1).bytes(n)                  # Adler32Tests.cs:22-40 lengths and beyond
    for seed in (1, 0, 0x00AB00CD):
        assert engine.adler32(d, seed) == zlib.adler32(d, seed)


def test_adler32_empty_returns_seed(engine):
    from zlibstream_b200 import Adler32
    for seed in (0, 1, 2):
        assert Adler32.Calculate(b"", seed, engine) == seed  # Adler32Tests.cs:13-20


def test_adler32_worst_case_bytes(engine):
    d = b"\xff" * (3 * 65536 + 5551)
    assert engine.adler32(d) == zlib.adler32(d)


@pytest.mark.parametrize("level,strategy,ss", [(1, 0, 65536), (6, 0, 65536), (9, 0, 262144), (6, 1, 65536), (6, 2, 65536),
                                               (1, 4, 65536), (3, 0, 131072), (6, 3, 65536)])
def test_inflate_reference_streams(engine, oracle, corpus, level, strategy, ss):
    """Streams produced by the oracle (= the reference's bytes): decoded bytes and Adler-32 must be exact."""
    data = b"".join(corpus[k] for k in sorted(corpus))
    comp, ends, adler = oracle.deflate_segments(data, ss, level, strategy)
    out, ad = engine.inflate_segments(comp, ends, ss, len(data))
    assert out == data and ad == adler == zlib.adler32(data)


def test_inflate_rle_carry_and_mixed(engine, oracle):
    """Q14: Rle segments may open with a run of the previous segment's last byte (whole segments of it for zeros);
    mixed-entropy tiles exercise stored blocks, long runs and long codes side by side in one warp."""
    for data, level, strategy, ss in ((bytes(400000), 6, 3, 65536), (corpus_data.png_like(700000, seed=3), 6, 3, 65536),
                                      (corpus_data.mixed_entropy(5 * 262144 + 77, seed=2), 6, 0, 262144),
                                      (corpus_data.mixed_entropy(9 * 65536 + 5, seed=4), 1, 0, 65536)):
        comp, ends, adler = oracle.deflate_segments(data, ss, level, strategy)
        out, ad = engine.inflate_segments(comp, ends, ss, len(data))
        assert out == data and ad == adler == zlib.adler32(data)


@pytest.mark.parametrize("level,strategy,ss", [(1, 0, 65536), (6, 0, 65536), (9, 0, 262144), (6, 2, 65536), (6, 3, 65536)])
def test_inflate_without_side_car(engine, oracle, corpus, level, strategy, ss):
    """a stream written by the reference has no segment offsets; the index is rebuilt from the full-flush
    markers (and must reproduce the reference's segment end offsets, i.e. ZlibOutputStream.TotalOut after each Write)."""
    data = b"".join(corpus[k] for k in sorted(corpus))
    comp, ends, adler = oracle.deflate_segments(data, ss, level, strategy)
    got_ends, total = engine.build_segment_index(comp, ss)
    assert got_ends == ends and total == len(data)
    out, ad = engine.inflate_segments(comp, None, ss, len(data))
    assert out == data and ad == adler


def test_index_scan_tolerates_chance_markers(engine, oracle):
    """00 00 FF FF inside stored blocks (verbatim input) and next to real markers must not split a segment."""
    rng = np.random.RandomState(11)
    raw = bytearray(rng.bytes(4 * 65536 + 999))
    for pos in (5, 1000, 65530, 65534, 65536, 70000, 131072 - 4, 200000, 262144 + 10, len(raw) - 4):
        raw[pos:pos + 4] = b"\x00\x00\xff\xff"
    text = corpus_data.text_like(3 * 65536, seed=5)
    for data in (bytes(raw), text[:65536] + bytes(raw[:65536]) + text[65536:] + b"\x00\x00\xff\xff" * 64):
        comp, ends, adler = oracle.deflate_segments(data, 65536, 6)
        assert comp.count(b"\x00\x00\xff\xff") > len(ends)
        got_ends, total = engine.build_segment_index(comp, 65536)
        assert got_ends == ends and total == len(data)
        out, ad = engine.inflate_segments(comp, None, 65536, len(data))
        assert out == data and ad == adler


def test_index_scan_on_damaged_streams(engine, oracle, corpus):
    """Without a side-car a damaged stream must end in the reference's exception (never a hang or a crash): bit flips in
    Huffman data, a destroyed marker, a truncated stream."""
    import random
    from zlibstream_b200 import ZlibStreamException
    data = corpus["alice29.txt"] + bytes(40000) + corpus["kennedy.xls"][:200000]
    comp, ends, adler = oracle.deflate_segments(data, 65536, 6)
    rnd = random.Random(7)
    bad_cases = []
    for _ in range(6):
        b = bytearray(comp)
        for _ in range(3):
            b[rnd.randrange(2, len(b) - 6)] ^= 1 << rnd.randrange(8)
        bad_cases.append(bytes(b))
    b = bytearray(comp)
    b[ends[1] - 1] = 0x00                      # second segment's marker destroyed: 00 00 FF 00
    bad_cases.append(bytes(b))
    bad_cases.append(comp[:ends[2] - 2] + comp[-6:])   # cut inside a marker
    for bad in bad_cases:
        try:
            out, ad = engine.inflate_segments(bad, None, 65536, len(data))
        except ZlibStreamException as e:
            assert str(e).startswith("inflating: ")
        else:
            assert out == data and ad == adler          # the flips happened to be harmless is not possible: Adler checks
    # and the engine is still usable afterwards
    out, ad = engine.inflate_segments(comp, None, 65536, len(data))
    assert out == data and ad == adler


def test_input_stream_without_offsets(engine, oracle, corpus):
    from zlibstream_b200 import ZlibInputStream
    data = corpus["alice29.txt"] + corpus["kennedy.xls"]
    comp, ends, adler = oracle.deflate_segments(data, 65536, 6)
    zi = ZlibInputStream(io.BytesIO(comp), segment_size=65536, engine=engine)
    assert zi.read() == data and zi.Adler == adler


def test_inflate_stored_and_degenerate(engine, oracle):
    rng = np.random.RandomState(3)
    for data in (rng.bytes(300001), bytes(400000), b"a", corpus_data.sparse_image(256, 300)):
        comp, ends, adler = oracle.deflate_segments(data, 65536, 6)
        out, ad = engine.inflate_segments(comp, ends, 65536, len(data))
        assert out == data and ad == adler


def test_inflate_detects_corruption(engine, oracle, corpus):
    from zlibstream_b200 import ZlibStreamException
    data = corpus["alice29.txt"]
    comp, ends, adler = oracle.deflate_segments(data, 65536, 6)
    bad = bytearray(comp)
    bad[-1] ^= 1
    with pytest.raises(ZlibStreamException, match="inflating: incorrect data check"):     # Inflate.cs:339
        engine.inflate_segments(bytes(bad), ends, 65536, len(data))
    bad = bytearray(comp)
    bad[0] = 0x79
    with pytest.raises(ZlibStreamException, match="inflating: unknown compression method"):  # Inflate.cs:134
        engine.inflate_segments(bytes(bad), ends, 65536, len(data))
    bad = bytearray(comp)
    bad[1] ^= 0x01
    with pytest.raises(ZlibStreamException, match="inflating: incorrect header check"):     # Inflate.cs:166
        engine.inflate_segments(bytes(bad), ends, 65536, len(data))
    bad = bytearray(comp)
    for i in range(3000, 3040):
        bad[i] ^= 0x5A
    with pytest.raises(ZlibStreamException, match="inflating: "):
        engine.inflate_segments(bytes(bad), ends, 65536, len(data))


def test_stream_classes_round_trip(engine, oracle, corpus):
    """ZlibOutputStream / ZlibInputStream with FullFlush: one Write per segment (ZlibOutputStream.cs:146), Dispose
    finishes; TotalOut and the bytes are the reference's."""
    from zlibstream_b200 import CompressionLevel, FlushMode, ZlibInputStream, ZlibOptions, ZlibOutputStream
    data = b"".join(corpus[k] for k in sorted(corpus))
    ss = 65536
    want, want_ends, _ = oracle.deflate_segments(data, ss, 6)
    sink = io.BytesIO()
    opts = ZlibOptions(CompressionLevel=CompressionLevel.Level6, FlushMode=FlushMode.FullFlush)
    z = ZlibOutputStream(sink, opts, engine=engine, batch_segments=16)
    for off in range(0, len(data), ss):
        z.Write(data, off, min(ss, len(data) - off))
    z.Flush()
    assert z.TotalIn == len(data) and z.TotalOut == want_ends[-1] and z.SegmentEndOffsets == want_ends
    z.close()
    assert sink.getvalue() == want and z.TotalOut == len(want)
    zi = ZlibInputStream(io.BytesIO(want), segment_size=ss, segment_end_offsets=want_ends, uncompressed_size=len(data),
                         engine=engine)
    got = bytearray(len(data))
    pos = 0
    while pos < len(data):
        r = zi.Read(got, pos, 8192)                         # ZlibStreamTests.Roundtrip.cs:100-103 read pattern
        if r == 0:
            break
        pos += r
    assert bytes(got) == data and zi.Adler == zlib.adler32(data)


def test_stream_classes_other_flush_modes(engine, oracle, corpus):
    """FlushMode.NoFlush (the options' default, ZlibOptions.cs:14-24) = one Write + Dispose: the reference's published size;
    FlushMode.SyncFlush: the oracle's stream for the same Writes; both read back as one piece.  Bounded read-ahead without a
    side-car for a full-flush stream."""
    from zlibstream_b200 import CompressionLevel, FlushMode, ZlibInputStream, ZlibOptions, ZlibOutputStream
    data = corpus["alice29.txt"]
    sink = io.BytesIO()
    z = ZlibOutputStream(sink, ZlibOptions(CompressionLevel=CompressionLevel.Level6), engine=engine)
    z.Write(data, 0, len(data))
    z.close()
    assert len(sink.getvalue()) == 55818 and sink.getvalue() == oracle.deflate_oneshot(data, 6)[0]
    zi = ZlibInputStream(io.BytesIO(sink.getvalue()), segment_size=None, uncompressed_size=len(data), engine=engine)
    assert zi.read() == data and zi.Adler == zlib.adler32(data)
    big = b"".join(corpus[k] for k in sorted(corpus))
    sink = io.BytesIO()
    z = ZlibOutputStream(sink, ZlibOptions(CompressionLevel=CompressionLevel.Level6, FlushMode=FlushMode.SyncFlush), engine=engine)
    for off in range(0, len(big), 65536):
        z.Write(big, off, min(65536, len(big) - off))
    z.close()
    want, want_ends = oracle.deflate_writes(big, 65536, 6, 0, 0, flush_mode=2)
    assert sink.getvalue() == want and z.SegmentEndOffsets == want_ends
    zi = ZlibInputStream(io.BytesIO(want), segment_size=None, uncompressed_size=len(big), engine=engine)
    assert zi.read() == big
    full, ends, adler = oracle.deflate_segments(big, 65536, 1)
    zi = ZlibInputStream(io.BytesIO(full), segment_size=65536, engine=engine, batch_segments=8)   # several read-ahead batches
    assert zi.read() == big and zi.Adler == adler and zi.TotalIn == len(full)


@pytest.mark.parametrize("level,strategy", [(1, 0), (6, 3)])
def test_piped_host_inflate_equals_the_input(engine, oracle, corpus, monkeypatch, level, strategy):
    """Host buffers + side-car + many segments: uploads, kernels and downloads overlap over 2-4 batches (the threshold is
    lowered here).  Rle: carried runs (Q14) cross the batch boundaries and the completed output is copied again."""
    monkeypatch.setenv("ZSB_INFLATE_PIPE_SEGMENTS", "16")
    ss = 65536
    data = (corpus["ptt5"] + bytes(3 * ss) + corpus["alice29.txt"]) * 3 + b"z" * 1000
    comp, ends, adler = oracle.deflate_segments_parallel(data, ss, level, strategy)
    assert len(ends) >= 32
    out, ad = engine.inflate_segments(comp, ends, ss, len(data))
    assert out == data and ad == adler
    bad = bytearray(comp)
    bad[ends[len(ends) // 2] - 40] ^= 0x55
    with pytest.raises(Exception, match="inflating: "):
        engine.inflate_segments(bytes(bad), ends, ss, len(data))


def test_stream_class_argument_behaviour(engine):
    from zlibstream_b200 import CompressionLevel, FlushMode, ZlibOptions, ZlibOutputStream, ZlibStreamException
    with pytest.raises(ValueError):
        ZlibOutputStream(None, CompressionLevel.Level6, engine=engine)
    z = ZlibOutputStream(io.BytesIO(), ZlibOptions(CompressionLevel=CompressionLevel.Level6, FlushMode=FlushMode.FullFlush),
                         engine=engine)
    z.write(b"x" * 65536)
    with pytest.raises(ZlibStreamException, match="deflating: "):   # equally sized writes only
        z.write(b"x" * 131072)
    with pytest.raises(io.UnsupportedOperation):
        z.seek(0)
    sink = io.BytesIO()
    z = ZlibOutputStream(sink, ZlibOptions(CompressionLevel=CompressionLevel.Level6, FlushMode=FlushMode.FullFlush), engine=engine)
    z.write(b"")
    z.close()
    assert sink.getvalue().hex() == "789c030000000001"      # empty stream: header, final empty static block, Adler(1)


def test_undersized_output_capacity_is_refused_without_overrun(engine, oracle, corpus):
    """out_cap two and a half segments short: every inflate kernel must stop at the caller's capacity (the bytes beyond
    it keep their guard pattern) and the call must fail with ZBUFERROR, not write a full segment past the buffer."""
    import torch
    from zlibstream_b200 import ZlibStreamException
    ss = 65536
    data = (corpus["lcet10.txt"] + corpus["plrabn12.txt"])[:9 * ss + 100]
    comp, ends, _ = oracle.deflate_segments(data, ss, 6)
    d_comp = torch.frombuffer(bytearray(comp), dtype=torch.uint8).cuda()
    d_ends = torch.tensor(ends, dtype=torch.int64, device="cuda")
    cap = len(data) - 2 * ss - ss // 2
    buf = torch.full((len(data) + ss,), 0xA5, dtype=torch.uint8, device="cuda")
    with pytest.raises(ZlibStreamException, match="inflating: "):
        engine.inflate_segments_device(d_comp.data_ptr(), len(comp), d_ends.data_ptr(), len(ends), ss, buf.data_ptr(), cap)
    assert engine._lib.zsb_last_error(engine._h) is not None
    assert bool((buf[cap:] == 0xA5).all())
    assert bytes(buf[:4 * ss].cpu().numpy()) == data[:4 * ss]      # the segments that fit are still decoded
    with pytest.raises(ZlibStreamException, match="inflating: "):  # host-buffer path: same refusal
        engine.inflate_segments(comp, ends, ss, cap)
