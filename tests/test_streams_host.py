"""Host logic of the stream classes (zlibstream_b200/streams.py) on CPU: batching, read-ahead, the marker cut and its retry,
trailer checks, flush modes, behaviour after an engine error.  The engine is a stand-in built on the oracle and stock zlib
(test infrastructure); the GPU tests drive the same classes with the real Engine."""
import io
import os
import sys
import zlib

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import corpus_data  # noqa: E402
import oracle_lib  # noqa: E402
from zlibstream_b200.enums import CompressionLevel, CompressionStrategy, FlushMode  # noqa: E402
from zlibstream_b200.options import ZlibOptions, ZlibStreamException  # noqa: E402
from zlibstream_b200.streams import MARKER, ZlibInputStream, ZlibOutputStream  # noqa: E402

SEG = 65536


class FakeEngine:
    def __init__(self):
        self.orc = oracle_lib.load()
        self.calls = []
        self.fail_next = False

    def deflate_segments(self, data, seg_size, level, strategy=0, hash_variant=0, first_segment_index=0, halo=None, header=True,
                         trailer=True, sync_flush=False, partial_flush=False):
        if self.fail_next:
            self.fail_next = False
            raise ZlibStreamException("deflating: device lost")
        data = bytes(data)
        self.calls.append(("deflate", len(data) // seg_size, first_segment_index, sync_flush))
        if sync_flush or partial_flush:
            out, ends = self.orc.deflate_writes(data, seg_size, level, strategy, hash_variant, flush_mode=2 if sync_flush else 1)
            return out, ends, zlib.adler32(data) & 0xFFFFFFFF
        if first_segment_index == 0:
            stream, ends, ad = self.orc.deflate_segments(data, seg_size, level, strategy, hash_variant)
            lo = 0 if header else 2
            return stream[lo:(len(stream) if trailer else ends[-1])], [e - lo for e in ends], ad
        prefix = bytes(halo[-1:]) * seg_size if halo else bytes(seg_size)   # only the byte before matters (Rle)
        stream, ends, _ = self.orc.deflate_segments(prefix + data, seg_size, level, strategy, hash_variant)
        return stream[ends[0]:ends[-1]], [e - ends[0] for e in ends[1:]], zlib.adler32(data) & 0xFFFFFFFF

    def deflate_stream(self, data, level, strategy=0, hash_variant=0):
        return self.orc.deflate_oneshot(bytes(data), level, strategy, hash_variant)

    def inflate_segments(self, comp, ends, seg_size, out_size, header=True, trailer=True):
        self.calls.append(("inflate", len(ends)))
        body = comp[2 if header else 0:ends[-1]]
        out = zlib.decompressobj(wbits=-15).decompress(body)
        assert len(out) <= out_size
        return out, zlib.adler32(out) & 0xFFFFFFFF

    def build_segment_index(self, comp, seg_size, header=True, trailer=True):
        start = 2 if header else 0
        end = len(comp) - (6 if trailer else 0)
        ends, total, pos = [], 0, start
        cand = start
        while True:
            cand = comp.find(MARKER, cand, end)
            if cand < 0:
                break
            e = cand + 4
            cand += 1
            try:
                d = zlib.decompressobj(wbits=-15)
                out = d.decompress(comp[pos:e])
            except zlib.error:
                continue
            if len(out) == seg_size or (e == end and 0 < len(out) <= seg_size):
                ends.append(e)
                total += len(out)
                pos = e
        if pos != end:
            raise ZlibStreamException("inflating: invalid stored block lengths")
        return ends, total

    def inflate_stream(self, comp, out_size):
        out = zlib.decompress(comp)
        return out, zlib.adler32(out) & 0xFFFFFFFF

    def close(self):
        pass


def _opts(level, mode, strategy=0):
    return ZlibOptions(CompressionLevel=CompressionLevel(level), CompressionStrategy=CompressionStrategy(strategy), FlushMode=mode)


def test_full_flush_batches_and_both_read_paths():
    eng = FakeEngine()
    data = corpus_data.text_like(10 * SEG + 4321, seed=4)
    want, want_ends, adler = eng.orc.deflate_segments(data, SEG, 6)
    sink = io.BytesIO()
    z = ZlibOutputStream(sink, _opts(6, FlushMode.FullFlush), engine=eng, batch_segments=3)
    for off in range(0, len(data), SEG):
        z.Write(data, off, min(SEG, len(data) - off))
    assert [c for c in eng.calls if c[0] == "deflate"] == [("deflate", 3, 0, False), ("deflate", 3, 3, False), ("deflate", 3, 6, False)]
    z.close()
    assert sink.getvalue() == want and z.SegmentEndOffsets == want_ends and z.TotalOut == len(want) and z.TotalIn == len(data)
    # with the side-car: four segments at a time
    eng.calls.clear()
    zi = ZlibInputStream(io.BytesIO(want), segment_size=SEG, segment_end_offsets=want_ends, uncompressed_size=len(data), engine=eng,
                         batch_segments=4)
    first = zi.read(1000)
    assert first == data[:1000] and eng.calls == [("inflate", 4)]          # only the first batch has been decoded
    rest = zi.read()
    assert first + rest == data and zi.Adler == adler and zi.TotalOut == len(data) and zi.TotalIn == len(want)
    assert [c[1] for c in eng.calls] == [4, 4, 3]
    # without it: batches cut at full-flush markers
    eng.calls.clear()
    zi = ZlibInputStream(io.BytesIO(want), segment_size=SEG, engine=eng, batch_segments=4)
    assert zi.read() == data and zi.Adler == adler
    assert len(eng.calls) >= 2 and sum(c[1] for c in eng.calls) == len(want_ends)


def test_no_flush_is_one_write_and_one_piece():
    eng = FakeEngine()
    data = corpus_data.load_corpus()["alice29.txt"]
    sink = io.BytesIO()
    z = ZlibOutputStream(sink, _opts(6, FlushMode.NoFlush), engine=eng)
    for off in range(0, len(data), 5000):
        z.write(data[off:off + 5000])
    z.close()
    assert len(sink.getvalue()) == 55818                                     # the reference's published size (benchmarks.md)
    assert sink.getvalue() == eng.orc.deflate_oneshot(data, 6)[0]
    zi = ZlibInputStream(io.BytesIO(sink.getvalue()), segment_size=None, uncompressed_size=len(data), engine=eng)
    assert zi.read() == data


def test_sync_flush_goes_out_in_one_call():
    eng = FakeEngine()
    data = corpus_data.text_like(4 * SEG + 99, seed=9)
    sink = io.BytesIO()
    z = ZlibOutputStream(sink, _opts(6, FlushMode.SyncFlush), engine=eng)
    for off in range(0, len(data), SEG):
        z.write(data[off:off + SEG])
    z.Flush()
    assert sink.getvalue() == b""                                            # nothing can leave before the stream is whole
    z.close()
    want, want_ends = eng.orc.deflate_writes(data, SEG, 6, 0, 0, flush_mode=2)
    assert sink.getvalue() == want and z.SegmentEndOffsets == want_ends and zlib.decompress(want) == data


def test_partial_flush_and_level0_trailer():
    eng = FakeEngine()
    data = corpus_data.text_like(3 * SEG + 5, seed=8)
    sink = io.BytesIO()
    z = ZlibOutputStream(sink, _opts(6, FlushMode.PartialFlush), engine=eng)
    for off in range(0, len(data), SEG):
        z.write(data[off:off + SEG])
    z.close()
    want, want_ends = eng.orc.deflate_writes(data, SEG, 6, 0, 0, flush_mode=1)
    assert sink.getvalue() == want and z.SegmentEndOffsets == want_ends and zlib.decompress(want) == data
    sink = io.BytesIO()
    z = ZlibOutputStream(sink, _opts(0, FlushMode.FullFlush), engine=eng)
    data = corpus_data.text_like(2 * SEG, seed=2)
    z.write(data[:SEG])
    z.write(data[SEG:])
    z.close()
    assert sink.getvalue() == eng.orc.deflate_segments(data, SEG, 0)[0] and sink.getvalue()[-9:-4] == b"\x01\x00\x00\xff\xff"
    zi = ZlibInputStream(io.BytesIO(sink.getvalue()), segment_size=SEG, engine=eng)
    assert zi.read() == data


def test_after_an_engine_error_nothing_more_is_written():
    eng = FakeEngine()
    sink = io.BytesIO()
    z = ZlibOutputStream(sink, _opts(1, FlushMode.FullFlush), engine=eng, batch_segments=2)
    z.write(b"a" * SEG)
    eng.fail_next = True
    with pytest.raises(ZlibStreamException, match="device lost"):
        z.write(b"b" * SEG)
    z.flush()                              # no second attempt, no exception
    with pytest.raises(ZlibStreamException, match="stream error"):
        z.write(b"c" * SEG)
    z.close()
    assert sink.getvalue() == b"" and not any(c[0] == "deflate" for c in eng.calls)


def test_reader_checks_the_trailer():
    eng = FakeEngine()
    data = corpus_data.text_like(3 * SEG, seed=1)
    comp, ends, _ = eng.orc.deflate_segments(data, SEG, 1)
    bad = bytearray(comp)
    bad[-1] ^= 1
    zi = ZlibInputStream(io.BytesIO(bytes(bad)), segment_size=SEG, segment_end_offsets=ends, engine=eng)
    with pytest.raises(ZlibStreamException, match="inflating: incorrect data check"):
        zi.read()
    zi = ZlibInputStream(io.BytesIO(comp[:-7] + b"\x07" + comp[-6:]), segment_size=SEG, segment_end_offsets=ends[:-1] + [ends[-1]], engine=eng)
    empty = ZlibInputStream(io.BytesIO(bytes.fromhex("789c030000000001")), segment_size=SEG, segment_end_offsets=[], engine=eng)
    assert empty.read() == b""
