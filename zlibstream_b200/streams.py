"""ZlibOutputStream / ZlibInputStream: host-side mirror of the reference's public stream classes.

Same names, arguments and error behaviour as ZlibOutputStream.cs / ZlibInputStream.cs; the engine underneath is
libzsb200 on a B200 instead of the managed Deflate / Inflate classes.  One `Write` with
`Options.FlushMode == FullFlush` is one segment (ZlibOutputStream.cs:146); because a single segment gives a GPU
nothing to parallelise, segments are queued and submitted in batches (SURVEY section 8b "host-side batching"):
compressed bytes reach BaseStream in order when a batch completes, on Flush() and on Dispose().

Only what the device path implements is accepted -- there is no CPU fallback:
  * compression needs FlushMode.FullFlush and equally sized writes (a multiple of 32 KiB, >= 64 KiB; the last write
    may be shorter);
  * decompression needs the segment size and, until the marker scan lands, the segment end offsets.
"""
import io
from typing import List, Optional, Sequence

from .engine import Engine, adler32_combine
from .enums import CompressionLevel, CompressionState, FlushMode, HashVariant
from .options import ZlibOptions, ZlibStreamException


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
        self._engine = engine or Engine(0)
        self._own_engine = engine is None
        self._batch = max(1, int(batch_segments))
        self._hash = hash_variant
        self._queue: List[bytes] = []
        self._seg_size: Optional[int] = None
        self._short_seen = False
        self._submitted = 0          # segments already compressed
        self._total_in = 0
        self._total_out = 0
        self._pending_in = 0
        self._adler = 1
        self._last_tail = None       # last 32 KiB submitted (halo for the next batch)
        self._finished = False
        self.SegmentEndOffsets: List[int] = []   # TotalOut after each Write, reference-exact

    # -- properties of the reference ---------------------------------------------------------------
    @property
    def TotalIn(self) -> int:
        return self._total_in + self._pending_in

    @property
    def TotalOut(self) -> int:
        """Exact once the queued batch has been submitted (Flush / Dispose / batch boundary)."""
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
        data = bytes(b)
        if not data:
            return 0  # ZlibOutputStream.cs:127-130
        if self._finished:
            raise ZlibStreamException("deflating: stream error")
        if self.Options.FlushMode != FlushMode.FullFlush:
            raise ZlibStreamException("deflating: the device path compresses full-flush segments only "
                                      "(Options.FlushMode = FlushMode.FullFlush); no CPU fallback exists")
        if self._short_seen:
            raise ZlibStreamException("deflating: only the last Write may be shorter than the segment size")
        if self._seg_size is None:
            n = len(data)
            if n < 65536 or n % 32768:
                # a single short write is allowed if it is the only one: size it as one 64 KiB-class segment
                self._seg_size = max(65536, (n + 32767) // 32768 * 32768)
            else:
                self._seg_size = n
        if len(data) > self._seg_size:
            raise ZlibStreamException("deflating: Write larger than the segment size")
        if len(data) < self._seg_size:
            self._short_seen = True
        self._queue.append(data)
        self._pending_in += len(data)
        if len(self._queue) >= self._batch:
            self._submit()
        return len(data)

    def _submit(self):
        if not self._queue:
            return
        data = b"".join(self._queue)
        first = self._submitted
        comp, ends, adler = self._engine.deflate_segments(
            data, self._seg_size, int(self.Options.CompressionLevel), int(self.Options.CompressionStrategy),
            int(self._hash), first_segment_index=first, halo=self._last_tail if first else None,
            header=(first == 0), trailer=False)
        self.BaseStream.write(comp)
        base = self._total_out
        self.SegmentEndOffsets.extend(base + e for e in ends)
        self._total_out = base + len(comp)
        self._adler = adler32_combine(self._adler, adler, len(data))
        self._total_in += len(data)
        self._pending_in = 0
        self._submitted += len(self._queue)
        tail = (self._last_tail or b"") + data
        self._last_tail = tail[-32768:] if len(tail) >= 32768 else None
        self._queue.clear()

    def Flush(self):
        self.flush()

    def flush(self):
        if self._finished or self.closed:
            return
        self._submit()
        if hasattr(self.BaseStream, "flush"):
            self.BaseStream.flush()  # ZlibOutputStream.cs:171

    def _finish(self):
        # ZlibOutputStream.Finish (ZlibOutputStream.cs:213-256): final empty static block + big-endian Adler-32
        if self._finished:
            return
        self._submit()
        tail = b""
        if self._submitted == 0:
            level = int(self.Options.CompressionLevel)
            level = 6 if level == -1 else level
            flevel = min(3, ((level - 1) & 0xFF) >> 1)  # Deflate.cs:466-480
            hdr = (0x78 << 8) | (flevel << 6)
            hdr += 31 - hdr % 31
            tail += bytes([hdr >> 8, hdr & 0xFF])
        from .dist import final_block
        tail += final_block(int(self.Options.CompressionLevel)) + self._adler.to_bytes(4, "big")
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
                super().close()


class ZlibInputStream(io.RawIOBase):
    """ZlibInputStream.cs.  Reads a full-flush-segmented zlib stream and serves decoded bytes."""

    def __init__(self, input, *, segment_size: int, segment_end_offsets: Optional[Sequence[int]] = None,
                 uncompressed_size: Optional[int] = None, engine: Optional[Engine] = None):
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
        self._decoded: Optional[bytes] = None
        self._pos = 0
        self._total_in = 0
        self.Adler = 1

    @property
    def TotalIn(self) -> int:
        return self._total_in

    @property
    def TotalOut(self) -> int:
        return self._pos

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

    def _decode_all(self):
        comp = self.BaseStream.read()
        self._total_in = len(comp)
        if self._ends is None:
            # a stream written by the reference carries no side-car: find the segments by their full-flush markers
            self._ends, total = self._engine.build_segment_index(comp, self._seg_size)
            if self._usize is None:
                self._usize = total
        nseg = len(self._ends)
        cap = self._usize if self._usize is not None else nseg * self._seg_size
        self._decoded, self.Adler = self._engine.inflate_segments(comp, self._ends, self._seg_size, cap)

    def Read(self, buffer, offset: int, count: int) -> int:
        mv = memoryview(buffer).cast("B")
        return self.readinto(mv[offset:offset + count])

    def ReadByte(self) -> int:
        b = self.read(1)
        return b[0] if b else -1

    def readinto(self, b) -> int:
        if self._decoded is None:
            self._decode_all()
        mv = memoryview(b).cast("B")
        n = min(len(mv), len(self._decoded) - self._pos)
        mv[:n] = self._decoded[self._pos:self._pos + n]
        self._pos += n
        return n

    def close(self):
        if not self.closed:
            if self._own_engine:
                self._engine.close()
            super().close()
