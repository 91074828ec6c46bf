"""Thin object wrapper over the C ABI: one `Engine` = one zsb_ctx on one GPU.

Plays the role of the reference's internal `ZLibStream` (ZlibStream.cs): the stream classes in streams.py
drive it exactly where the reference's ZlibOutputStream / ZlibInputStream drive ZLibStream.Deflate / Inflate.
"""
import ctypes
from typing import Optional, Sequence, Tuple

from . import _native
from .enums import CompressionState, HashVariant
from .options import ZlibStreamException


def _addr(buf) -> Tuple[int, int, object]:
    """(address, nbytes, keepalive) of a bytes-like host buffer."""
    if isinstance(buf, (bytes, bytearray)):
        n = len(buf)
        if isinstance(buf, bytes):
            keep = ctypes.c_char_p(buf)
            return ctypes.cast(keep, ctypes.c_void_p).value or 0, n, (keep, buf)
        arr = (ctypes.c_char * n).from_buffer(buf) if n else None
        return (ctypes.addressof(arr) if n else 0), n, arr
    mv = memoryview(buf).cast("B")
    n = mv.nbytes
    if n == 0:
        return 0, 0, mv
    if mv.readonly:
        b = bytes(mv)
        keep = ctypes.c_char_p(b)
        return ctypes.cast(keep, ctypes.c_void_p).value, n, (keep, b)
    arr = (ctypes.c_char * n).from_buffer(mv)
    return ctypes.addressof(arr), n, (arr, mv)


class Engine:
    def __init__(self, device: int = 0):
        self._lib = _native.load()
        h = ctypes.c_void_p()
        rc = self._lib.zsb_ctx_create(device, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(f"zsb_ctx_create(device={device}) failed with {CompressionState(rc).name}: "
                               "libzsb200 needs a CUDA device of compute capability >= 10.0 (no CPU fallback)")
        self._h = h
        self.device = device

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.zsb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- helpers ----------------------------------------------------------------------------------
    @property
    def message(self) -> str:
        return self._lib.zsb_last_error(self._h).decode()

    @property
    def launch_count(self) -> int:
        return int(self._lib.zsb_ctx_launch_count(self._h))

    def set_stream(self, cuda_stream: int):
        self._lib.zsb_ctx_set_stream(self._h, ctypes.c_void_p(cuda_stream))

    def last_kernel_ms(self):
        a, b, c = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        self._lib.zsb_ctx_last_kernel_ms(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        d = (ctypes.c_float * 5)()
        self._lib.zsb_ctx_last_kernel_ms_detail(self._h, d, 5)
        return {"lz77": a.value, "huffman": b.value, "inflate": c.value, "lz77_lists": d[3], "lz77_resolve": d[4]}

    def deflate_bound(self, n: int, seg_size: int) -> int:
        return int(self._lib.zsb_deflate_bound(n, seg_size))

    def _check(self, rc: int, compressing: bool):
        if rc not in (CompressionState.ZOK, CompressionState.ZSTREAMEND):
            # ZlibOutputStream.cs:149-152 / ThrowHelper.cs:21-23
            raise ZlibStreamException(("de" if compressing else "in") + "flating: " + self.message)

    # -- host-buffer API --------------------------------------------------------------------------
    def deflate_segments(self, data, seg_size: int, level: int, strategy: int = 0,
                         hash_variant: int = HashVariant.Crc32C, first_segment_index: int = 0, halo=None,
                         header: bool = True, trailer: bool = True, sync_flush: bool = False, partial_flush: bool = False):
        """Returns (compressed bytes, list of segment end offsets, adler32 of data).  sync_flush: every Write ends with
        FlushMode.SyncFlush instead of FullFlush (segments share the hash table: serial on the device, one-piece to inflate)."""
        addr, n, keep = _addr(data)
        nseg = (n + seg_size - 1) // seg_size if seg_size else 0
        cap = self.deflate_bound(n, seg_size)
        out = ctypes.create_string_buffer(cap)
        out_len = ctypes.c_size_t(0)
        ends = (ctypes.c_uint64 * max(nseg, 1))()
        adler = ctypes.c_uint32(0)
        flags = (0 if header else _native.ZSB_NO_HEADER) | (0 if trailer else _native.ZSB_NO_TRAILER)
        flags |= _native.ZSB_SYNC_FLUSH if sync_flush else 0
        flags |= _native.ZSB_PARTIAL_FLUSH if partial_flush else 0
        haddr, hkeep = 0, None
        if halo is not None:
            haddr, hn, hkeep = _addr(halo)
            if hn != 32768:
                raise ValueError("halo must be the 32768 bytes preceding the data")
        rc = self._lib.zsb_deflate_segments(self._h, addr, n, seg_size, int(level), int(strategy), int(hash_variant),
                                            first_segment_index, haddr, flags, out, cap, ctypes.byref(out_len),
                                            ctypes.cast(ends, ctypes.c_void_p), ctypes.byref(adler))
        self._check(rc, True)
        del keep, hkeep
        return out.raw[:out_len.value], list(ends[:nseg]), adler.value

    def inflate_segments(self, comp, seg_end_offsets: Optional[Sequence[int]], seg_size: int, out_size: int,
                         header: bool = True, trailer: bool = True):
        """Returns (decoded bytes, adler32)."""
        addr, n, keep = _addr(comp)
        nseg = len(seg_end_offsets) if seg_end_offsets is not None else 0
        ends = (ctypes.c_uint64 * max(nseg, 1))(*(seg_end_offsets or []))
        out = ctypes.create_string_buffer(max(out_size, 1))
        out_len = ctypes.c_size_t(0)
        adler = ctypes.c_uint32(0)
        flags = (0 if header else _native.ZSB_NO_HEADER) | (0 if trailer else _native.ZSB_NO_TRAILER)
        rc = self._lib.zsb_inflate_segments(self._h, addr, n,
                                            ctypes.cast(ends, ctypes.c_void_p) if seg_end_offsets is not None else None,
                                            nseg, seg_size, flags, out, out_size, ctypes.byref(out_len),
                                            ctypes.byref(adler))
        self._check(rc, False)
        del keep
        return out.raw[:out_len.value], adler.value

    def deflate_stream(self, data, level: int, strategy: int = 0, hash_variant: int = HashVariant.Crc32C):
        """One Write with FlushMode.NoFlush + Dispose (the reference's benchmark call pattern, DeflateCorpusBenchmark.cs:86-99):
        a one-piece zlib stream, at most 64 MiB of input.  Returns (stream bytes, adler32 of data)."""
        addr, n, keep = _addr(data)
        cap = self.deflate_bound(n, 65536) + 64
        out = ctypes.create_string_buffer(cap)
        out_len = ctypes.c_size_t(0)
        adler = ctypes.c_uint32(0)
        rc = self._lib.zsb_deflate_segments(self._h, addr, n, 0, int(level), int(strategy), int(hash_variant), 0, None,
                                            _native.ZSB_SINGLE_STREAM, out, cap, ctypes.byref(out_len), None, ctypes.byref(adler))
        self._check(rc, True)
        del keep
        return out.raw[:out_len.value], adler.value

    def inflate_stream(self, comp, out_size: int):
        """Decodes a one-piece zlib stream (any conforming deflater's, e.g. the reference's with FlushMode.NoFlush).
        Returns (decoded bytes, adler32); raises "incorrect data check" if the trailer does not match."""
        addr, n, keep = _addr(comp)
        out = ctypes.create_string_buffer(max(out_size, 1))
        out_len = ctypes.c_size_t(0)
        adler = ctypes.c_uint32(0)
        rc = self._lib.zsb_inflate_segments(self._h, addr, n, None, 0, 0, _native.ZSB_SINGLE_STREAM, out, out_size,
                                            ctypes.byref(out_len), ctypes.byref(adler))
        self._check(rc, False)
        del keep
        return out.raw[:out_len.value], adler.value

    def build_segment_index(self, comp, seg_size: int, header: bool = True, trailer: bool = True):
        """Segment index of a full-flush-segmented stream without a side-car (zsb_build_segment_index).
        Returns ([segment end offsets], uncompressed size)."""
        addr, n, keep = _addr(comp)
        flags = (0 if header else _native.ZSB_NO_HEADER) | (0 if trailer else _native.ZSB_NO_TRAILER)
        nseg = ctypes.c_size_t(0)
        total = ctypes.c_uint64(0)
        # one call scans, validates and fills; 64 KiB of anything compresses to more than 64 bytes, so n / 64 slots nearly always
        # suffice (the library reports the count with ZBUFERROR if they do not, and the call is repeated once)
        cap = n // 64 + 16
        ends = (ctypes.c_uint64 * cap)()
        rc = self._lib.zsb_build_segment_index(self._h, addr, n, seg_size, flags, ends, cap, ctypes.byref(nseg), ctypes.byref(total))
        if rc == CompressionState.ZBUFERROR and nseg.value > cap:
            cap = nseg.value
            ends = (ctypes.c_uint64 * cap)()
            rc = self._lib.zsb_build_segment_index(self._h, addr, n, seg_size, flags, ends, cap, ctypes.byref(nseg), ctypes.byref(total))
        self._check(rc, False)
        del keep
        return list(ends)[:nseg.value], total.value

    def png_encode(self, pixels, width: int, height: int, color_type: int = 6, filter: int = 5, level: int = 6,
                   strategy: int = 1, seg_size: int = 65536, idat_bytes: int = 0) -> bytes:
        """A complete PNG file from 8-bit samples (host buffer, row-major, no padding): filtered, deflated, check-summed and
        framed on the device (zsb_png_encode).  filter 0..4 = PNG filter types, 5 = adaptive per scanline."""
        addr, n, keep = _addr(pixels)
        cap = int(self._lib.zsb_png_bound(width, height, color_type, seg_size, idat_bytes))
        if cap == 0:
            raise ValueError("color_type")
        out = ctypes.create_string_buffer(cap)
        out_len = ctypes.c_size_t(0)
        rc = self._lib.zsb_png_encode(self._h, addr, width, height, color_type, filter, int(level), int(strategy), seg_size,
                                      idat_bytes, 0, out, cap, ctypes.byref(out_len))
        self._check(rc, True)
        del keep
        return out.raw[:out_len.value]

    def png_encode_device(self, px_ptr: int, width: int, height: int, color_type: int, filter: int, level: int, strategy: int,
                          out_ptr: int, out_cap: int, seg_size: int = 65536, idat_bytes: int = 0) -> int:
        """Device pointers in and out; returns the PNG's length."""
        out_len = ctypes.c_size_t(0)
        rc = self._lib.zsb_png_encode(self._h, px_ptr, width, height, color_type, filter, int(level), int(strategy), seg_size,
                                      idat_bytes, _native.ZSB_IN_DEVICE | _native.ZSB_OUT_DEVICE, out_ptr, out_cap,
                                      ctypes.byref(out_len))
        self._check(rc, True)
        return out_len.value

    def adler32(self, data, seed: int = 1) -> int:
        addr, n, keep = _addr(data)
        res = ctypes.c_uint32(0)
        rc = self._lib.zsb_adler32(self._h, seed, addr, n, 0, ctypes.byref(res))
        self._check(rc, True)
        del keep
        return res.value

    # -- device-pointer API (used by bench.py / dist.py with torch tensors) ------------------------
    def deflate_segments_device(self, in_ptr: int, n: int, seg_size: int, level: int, strategy: int, hash_variant: int,
                                first_segment_index: int, out_ptr: int, out_cap: int, seg_end_ptr: int = 0,
                                header: bool = True, trailer: bool = True, halo_ptr: int = 0):
        """All pointers are device addresses. Returns (out_len, adler)."""
        out_len = ctypes.c_size_t(0)
        adler = ctypes.c_uint32(0)
        flags = _native.ZSB_IN_DEVICE | _native.ZSB_OUT_DEVICE | _native.ZSB_OFFSETS_DEVICE
        flags |= (0 if header else _native.ZSB_NO_HEADER) | (0 if trailer else _native.ZSB_NO_TRAILER)
        rc = self._lib.zsb_deflate_segments(self._h, in_ptr, n, seg_size, int(level), int(strategy), int(hash_variant),
                                            first_segment_index, halo_ptr or None, flags, out_ptr, out_cap,
                                            ctypes.byref(out_len), seg_end_ptr or None, ctypes.byref(adler))
        self._check(rc, True)
        return out_len.value, adler.value

    def inflate_segments_device(self, in_ptr: int, n: int, seg_end_ptr: int, nseg: int, seg_size: int, out_ptr: int,
                                out_cap: int, header: bool = True, trailer: bool = True, continuation: bool = False):
        """continuation: `in` is a later shard of a stream (its first segment may open with a Q14 run, see dist.py)."""
        out_len = ctypes.c_size_t(0)
        adler = ctypes.c_uint32(0)
        flags = _native.ZSB_IN_DEVICE | _native.ZSB_OUT_DEVICE | _native.ZSB_OFFSETS_DEVICE
        flags |= (0 if header else _native.ZSB_NO_HEADER) | (0 if trailer else _native.ZSB_NO_TRAILER)
        flags |= _native.ZSB_CONTINUATION if continuation else 0
        rc = self._lib.zsb_inflate_segments(self._h, in_ptr, n, seg_end_ptr, nseg, seg_size, flags, out_ptr, out_cap,
                                            ctypes.byref(out_len), ctypes.byref(adler))
        self._check(rc, False)
        return out_len.value, adler.value

    def last_inflate_carry(self) -> int:
        """Q14: leading bytes of the last inflate call's first segment that repeat the byte before the call's input."""
        c = ctypes.c_uint32(0)
        self._lib.zsb_ctx_last_inflate_carry(self._h, ctypes.byref(c))
        return c.value

    def inflate_fill_carry(self, prev_byte: int) -> int:
        """Q14: completes the device output of the last inflate call with the byte before it; returns its Adler-32."""
        a = ctypes.c_uint32(0)
        rc = self._lib.zsb_inflate_fill_carry(self._h, prev_byte & 0xFF, ctypes.byref(a))
        self._check(rc, False)
        return a.value

    def adler32_device(self, ptr: int, n: int, seed: int = 1) -> int:
        res = ctypes.c_uint32(0)
        rc = self._lib.zsb_adler32(self._h, seed, ptr, n, _native.ZSB_IN_DEVICE, ctypes.byref(res))
        self._check(rc, True)
        return res.value

    def debug_fetch(self, what: int, byte_offset: int, nbytes: int) -> bytes:
        buf = ctypes.create_string_buffer(nbytes)
        rc = self._lib.zsb_debug_fetch(self._h, what, byte_offset, buf, nbytes)
        if rc != 0:
            raise RuntimeError("zsb_debug_fetch: " + self.message)
        return buf.raw


def adler32_combine(adler1: int, adler2: int, len2: int) -> int:
    """adler32_combine (absent from the reference). Pure host arithmetic in libzsb200."""
    return int(_native.load().zsb_adler32_combine(adler1, adler2, len2))
