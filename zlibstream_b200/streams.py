"""ZlibOutputStream / ZlibInputStream: host-side mirror of the reference's public stream classes.

Same names, arguments and error behaviour as ZlibOutputStream.cs / ZlibInputStream.cs; the engine underneath is
libzsb200 on a B200 instead of the managed Deflate / Inflate classes.  There is no CPU fallback.

Writing.  One `Write` with `Options.FlushMode == FullFlush` is one segment (ZlibOutputStream.cs:146); because a single
segment gives a GPU nothing to parallelise, segments are staged in ONE buffer and submitted in batches (SURVEY section 8b
"host-side batching"): compressed bytes reach BaseStream in order when a batch completes, on Flush() and on Dispose().
  * FullFlush  equally sized writes (a multiple of 32 KiB, >= 64 KiB; the last write may be shorter), batched.
  * SyncFlush, PartialFlush  the same writes, but the hash table survives the flush (Deflate.cs:566-595; a partial flush
               ends with Tr_align instead of the stored marker and does not even align the stream to a byte): the segments of the
               whole stream go to the device in one call at Dispose and are parsed one after the other.
  * NoFlush    everything written is compressed as ONE Write at Dispose (the reference's benchmark call pattern,
               DeflateCorpusBenchmark.cs:86-99).  The reference's output depends on where its Write calls fall even without a
               flush, so the bytes equal the reference's for a caller that makes one Write; at most 64 MiB.

Reading.  ZlibInputStream reads ahead a bounded batch of compressed bytes from BaseStream, finds the segments in it (from
the side-car offsets, or by the full-flush markers: zsb_build_segment_index), decodes the batch on the device and serves
`Read` from it -- the reference reads 8192 bytes ahead and inflates as it goes (ZlibInputStream.cs:133-186).  A stream without
full flushes (`segment_size=None`) is one piece and is decoded in one call.
"""
import io
from typing import List, Optional, Sequence

from .dist import final_block
from .engine import Engine, adler32_combine
from .enums import CompressionLevel, FlushMode, HashVariant
from .options import ZlibOptions, ZlibStreamException

MARKER = b"\x00\x00\xff\xff"


class _NotSupported(io.UnsupportedOperation):
    pass


class ZlibOutputStream(io.RawIOBase):
    """ZlibOutputStream.cs.  Constructors: (output), (output, level), (output, options)."""

    def __init__(self, output, options=None, *, engine: Optional[Engine] = None, batch_segments: int = 4096,
                 hash_variant: HashVariant = HashVariant.Crc32C):
        super().__init__()
        if output is None:
            raise ValueError("output")  # ArgumentNullException
        if options is None:
            options = ZlibOptions()
        elif not isinstance(options, ZlibOptions):
            options = ZlibOptions(CompressionLevel=CompressionLevel(options))
        self.BaseStream = output
        self.Options = options
        self._compress = options.CompressionLevel is not None  # ZlibStream.cs:20-28
        if not self._compress:
            raise ZlibStreamException("inflating: a ZlibOutputStream without a CompressionLevel inflates in the "
                                      "reference; use ZlibInputStream with the device path")
        level = int(options.CompressionLevel)
        if level < -1 or level > 9:
            raise ValueError("level")  # ArgumentOutOfRangeException (Deflate.cs:273-276)
        if not 0 <= int(options.CompressionStrategy) <= 4:
            raise ValueError("strategy")  # Deflate.cs:278-281
        self._level = 6 if level == -1 else level
        self._engine = engine or Engine(0)
        self._own_engine = engine is None
        self._batch = max(1, int(batch_segments))
        self._hash = hash_variant
        self._stage = bytearray()    # the writes of the current batch, back to back
        self._staged_segments = 0
        self._seg_size: Optional[int] = None
        self._short_seen = False
        self._submitted = 0          # segments already compressed
        self._total_in = 0
        self._total_out = 0
        self._adler = 1
        self._last_tail = None       # last 32 KiB submitted (halo for the next batch)
        self._finished = False
        self._failed = False
        self.SegmentEndOffsets: List[int] = []   # TotalOut after each Write, reference-exact

    # -- properties of the reference ---------------------------------------------------------------
    @property
    def TotalIn(self) -> int:
        return self._total_in + len(self._stage)

    @property
    def TotalOut(self) -> int:
        """Exact once the staged batch has been submitted (Flush / Dispose / batch boundary)."""
        return self._total_out

    def readable(self):
        return False

    def seekable(self):
        return False

    def writable(self):
        return True

    def seek(self, *a):
        raise _NotSupported("Seek")  # NotSupportedException (ZlibOutputStream.cs:174-183)

    def truncate(self, *a):
        raise _NotSupported("SetLength")

    def tell(self):
        raise _NotSupported("Position")

    def readinto(self, b):
        raise _NotSupported("Read")

    # -- writing ---------------------------------------------------------------------------------
    def Write(self, buffer, offset: int = 0, count: Optional[int] = None):
        if buffer is None:
            raise ValueError("buffer")
        mv = memoryview(buffer).cast("B")
        if count is None:
            count = len(mv) - offset
        self.write(mv[offset:offset + count])

    def WriteByte(self, value: int):
        self.write(bytes([value & 0xFF]))

    def write(self, b) -> int:
        mv = memoryview(b).cast("B")
        n = len(mv)
        if not n:
            return 0  # ZlibOutputStream.cs:127-130
        if self._finished or self._failed:
            raise ZlibStreamException("deflating: stream error")
        mode = self.Options.FlushMode
        if mode == FlushMode.NoFlush:
            if len(self._stage) + n > (64 << 20):
                raise ZlibStreamException("deflating: a stream without flushes holds at most 64 MiB on the device path")
            self._stage += mv
            return n
        if self._short_seen:
            raise ZlibStreamException("deflating: only the last Write may be shorter than the segment size")
        if self._seg_size is None:
            # a single short write is allowed if it is the only one: size it as one 64 KiB-class segment
            self._seg_size = n if (n >= 65536 and n % 32768 == 0) else max(65536, (n + 32767) // 32768 * 32768)
        if n > self._seg_size:
            raise ZlibStreamException("deflating: Write larger than the segment size")
        if n < self._seg_size:
            self._short_seen = True
        self._stage += mv            # the one copy of the caller's bytes; the engine reads the staging buffer in place
        self._staged_segments += 1
        if mode == FlushMode.FullFlush and self._staged_segments >= self._batch:
            self._submit()
        return n

    def _emit(self, comp: bytes, ends, adler: int, nbytes: int):
        self.BaseStream.write(comp)
        base = self._total_out
        self.SegmentEndOffsets.extend(base + e for e in ends)
        self._total_out = base + len(comp)
        self._adler = adler32_combine(self._adler, adler, nbytes)
        self._total_in += nbytes

    def _submit(self):
        """FullFlush: compresses the staged segments (independent work) and hands the bytes to BaseStream."""
        if not self._stage or self._failed:
            return
        first = self._submitted
        try:
            comp, ends, adler = self._engine.deflate_segments(
                self._stage, self._seg_size, self._level, int(self.Options.CompressionStrategy), int(self._hash),
                first_segment_index=first, halo=self._last_tail if first else None, header=(first == 0), trailer=False)
        except Exception:
            self._failed = True      # nothing more is compressed or written: flush() and close() become no-ops
            raise
        self._emit(comp, ends, adler, len(self._stage))
        self._submitted += self._staged_segments
        tail = (bytes(self._last_tail) if self._last_tail else b"") + bytes(self._stage[-32768:])
        self._last_tail = tail[-32768:] if len(tail) >= 32768 else None
        self._stage = bytearray()
        self._staged_segments = 0

    def Flush(self):
        self.flush()

    def flush(self):
        if self._finished or self._failed or self.closed:
            return
        if self.Options.FlushMode == FlushMode.FullFlush:
            self._submit()
        if hasattr(self.BaseStream, "flush"):
            self.BaseStream.flush()  # ZlibOutputStream.cs:171

    def _finish(self):
        # ZlibOutputStream.Finish (ZlibOutputStream.cs:213-256)
        if self._finished or self._failed:
            return
        mode = self.Options.FlushMode
        strategy = int(self.Options.CompressionStrategy)
        try:
            if mode == FlushMode.NoFlush:
                comp, adler = self._engine.deflate_stream(self._stage, self._level, strategy, int(self._hash))
                self._emit(comp, [], adler, len(self._stage))
                self._stage = bytearray()
                self._finished = True
            elif mode in (FlushMode.SyncFlush, FlushMode.PartialFlush) and self._stage:
                comp, ends, adler = self._engine.deflate_segments(self._stage, self._seg_size, self._level, strategy,
                                                                  int(self._hash), sync_flush=mode == FlushMode.SyncFlush,
                                                                  partial_flush=mode == FlushMode.PartialFlush)
                self._emit(comp, ends, adler, len(self._stage))     # header, segments and trailer in one piece
                self._stage = bytearray()
                self._finished = True
        except Exception:
            self._failed = True
            raise
        if not self._finished:
            self._submit()
            tail = b""
            if self._submitted == 0:
                flevel = min(3, ((self._level - 1) & 0xFF) >> 1)  # Deflate.cs:466-480
                hdr = (0x78 << 8) | (flevel << 6)
                hdr += 31 - hdr % 31
                tail += bytes([hdr >> 8, hdr & 0xFF])
            tail += final_block(self._level) + self._adler.to_bytes(4, "big")
            self.BaseStream.write(tail)
            self._total_out += len(tail)
            self._finished = True
        if hasattr(self.BaseStream, "flush"):
            self.BaseStream.flush()

    def Dispose(self):
        self.close()

    def close(self):
        if not self.closed:
            try:
                self._finish()
            finally:
                if self._own_engine:
                    self._engine.close()
                self._finished = True    # io.RawIOBase.close() calls flush(): nothing is left to do, also after an error
                super().close()


class ZlibInputStream(io.RawIOBase):
    """ZlibInputStream.cs.  Reads a zlib stream and serves decoded bytes.

    segment_size          uncompressed size of the full-flush segments (the writer's Write size); None = the stream has
                          no full flushes (FlushMode.NoFlush or SyncFlush) and is decoded as one piece
    segment_end_offsets   optional side-car: the writer's TotalOut after each Write
    uncompressed_size     needed for a one-piece stream (its output buffer); optional otherwise
    batch_segments        how many segments are read ahead and decoded at a time (bounds memory)
    """

    def __init__(self, input, *, segment_size: Optional[int], segment_end_offsets: Optional[Sequence[int]] = None,
                 uncompressed_size: Optional[int] = None, engine: Optional[Engine] = None, batch_segments: int = 4096):
        super().__init__()
        if input is None:
            raise ValueError("input")
        self.BaseStream = input
        self.Options = ZlibOptions()  # CompressionLevel None => inflate (ZlibStream.cs:20-28)
        self._engine = engine or Engine(0)
        self._own_engine = engine is None
        self._seg_size = segment_size
        self._ends = list(segment_end_offsets) if segment_end_offsets is not None else None
        self._usize = uncompressed_size
        self._batch = max(1, int(batch_segments))
        self._buf = bytearray()      # compressed bytes read ahead, not yet decoded
        self._buf_offset = 0         # stream offset of _buf[0]
        self._next_seg = 0           # with a side-car: the next segment to decode
        self._decoded = b""          # the current decoded batch
        self._dpos = 0
        self._eof_in = False
        self._done = False
        self._first = True
        self._out_total = 0
        self._total_in = 0
        self.Adler = 1

    @property
    def TotalIn(self) -> int:
        return self._total_in

    @property
    def TotalOut(self) -> int:
        return self._out_total - (len(self._decoded) - self._dpos)

    def readable(self):
        return True

    def seekable(self):
        return False

    def writable(self):
        return False

    def seek(self, *a):
        raise _NotSupported("Seek")

    def tell(self):
        raise _NotSupported("Position")

    def write(self, b):
        raise _NotSupported("Write")

    # -- read-ahead ---------------------------------------------------------------------------------
    def _pull(self, want: int):
        """Reads until _buf holds `want` bytes or BaseStream ends."""
        while len(self._buf) < want and not self._eof_in:
            chunk = self.BaseStream.read(max(8192, min(want - len(self._buf), 1 << 26)))   # the reference reads 8192-byte chunks
            if not chunk:
                self._eof_in = True
                break
            self._buf += chunk
            self._total_in += len(chunk)

    def _finish_stream(self, tail: bytes):
        """`tail` = what follows the last segment: the final block and the big-endian Adler-32 (Inflate.cs:286-342)."""
        if tail[:2] == b"\x03\x00":
            want = tail[2:6]
        elif tail[:5] == b"\x01\x00\x00\xff\xff":
            want = tail[5:9]
        else:
            raise ZlibStreamException("inflating: invalid block type")
        if len(want) < 4:
            raise ZlibStreamException("inflating: buffer error")
        if int.from_bytes(want, "big") != self.Adler:
            raise ZlibStreamException("inflating: incorrect data check")
        self._done = True

    def _take(self, comp: bytes, ends_rel, nbytes_hint):
        """Decodes one batch: comp = [header] segments, ends_rel relative to comp."""
        nseg = len(ends_rel)
        cap = nseg * self._seg_size if nbytes_hint is None else nbytes_hint
        out, ad = self._engine.inflate_segments(comp, ends_rel, self._seg_size, cap, header=self._first, trailer=False)
        self.Adler = adler32_combine(self.Adler, ad, len(out))
        self._first = False
        self._decoded, self._dpos = out, 0
        self._out_total += len(out)

    def _next_batch(self):
        if self._done:
            return
        if self._seg_size is None:
            # one piece: the whole stream in one call (its final block ends it; Adler-32 checked by the engine)
            self._pull(1 << 62)
            if self._usize is None:
                raise ZlibStreamException("inflating: a one-piece stream needs uncompressed_size")
            out, ad = self._engine.inflate_stream(bytes(self._buf), self._usize)
            self._buf = bytearray()
            self.Adler = ad
            self._decoded, self._dpos = out, 0
            self._out_total += len(out)
            self._done = True
            return
        if self._ends is not None:
            k0, k1 = self._next_seg, min(len(self._ends), self._next_seg + self._batch)
            if k0 == k1:
                self._pull(9)
                hdr = 2 if self._first else 0       # an empty stream is its header and its trailer
                self._finish_stream(bytes(self._buf[hdr:]))
                return
            self._pull(self._ends[k1 - 1] - self._buf_offset)
            if len(self._buf) < self._ends[k1 - 1] - self._buf_offset:
                raise ZlibStreamException("inflating: buffer error")
            cut = self._ends[k1 - 1] - self._buf_offset
            rel = [e - self._buf_offset for e in self._ends[k0:k1]]
            hint = None
            if k1 == len(self._ends) and self._usize is not None:
                hint = self._usize - k0 * self._seg_size
            self._take(bytes(self._buf[:cut]), rel, hint)
            del self._buf[:cut]
            self._buf_offset += cut
            self._next_seg = k1
            return
        # no side-car: read ahead about one batch of compressed bytes and cut it at a full-flush marker; a chance
        # occurrence of the marker inside compressed data makes the index fail (its last piece does not end at a flush point),
        # and the batch is retried with more input
        want = max(1 << 16, self._batch * self._seg_size // 2)
        while True:
            self._pull(want)
            hdr = 2 if self._first else 0
            if self._eof_in:
                stored_final = bytes(self._buf[-9:-4]) == b"\x01\x00\x00\xff\xff" and bytes(self._buf[-6:-4]) != b"\x03\x00"
                tl = 9 if stored_final else 6
                if len(self._buf) <= hdr + tl:       # nothing but the trailer is left
                    self._finish_stream(bytes(self._buf[hdr:]))
                    return
                cut = len(self._buf) - tl
            else:
                cut = self._buf.rfind(MARKER, hdr, len(self._buf) - 9)
                if cut < 0:
                    want *= 2
                    continue
                cut += 4
            try:
                ends, total = self._engine.build_segment_index(bytes(self._buf[:cut]), self._seg_size, header=self._first, trailer=False)
            except ZlibStreamException:
                if self._eof_in:
                    raise
                want *= 2
                continue
            self._take(bytes(self._buf[:cut]), ends, total)
            del self._buf[:cut]
            self._buf_offset += cut
            return

    def Read(self, buffer, offset: int, count: int) -> int:
        mv = memoryview(buffer).cast("B")
        return self.readinto(mv[offset:offset + count])

    def ReadByte(self) -> int:
        b = self.read(1)
        return b[0] if b else -1

    def readinto(self, b) -> int:
        mv = memoryview(b).cast("B")
        while self._dpos >= len(self._decoded) and not self._done:
            self._next_batch()
        n = min(len(mv), len(self._decoded) - self._dpos)
        mv[:n] = self._decoded[self._dpos:self._dpos + n]
        self._dpos += n
        return n

    def close(self):
        if not self.closed:
            if self._own_engine:
                self._engine.close()
            super().close()
