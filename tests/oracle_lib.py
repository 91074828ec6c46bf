"""ctypes wrapper of the CPU oracle (oracle/libzso.so).  Test infrastructure only."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.environ.get("ZSO_LIB") or os.path.join(ORACLE_DIR, "libzso.so")   # ZSO_LIB: the ASan build (make -C oracle asan)

_lib = None


class BlockTrace(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("last", ctypes.c_int), ("n_symbols", ctypes.c_int),
                ("stored_len", ctypes.c_int), ("bit_start", ctypes.c_uint64)]


BLOCK_CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.POINTER(BlockTrace), ctypes.POINTER(ctypes.c_uint16),
                            ctypes.POINTER(ctypes.c_uint8))


class Counters(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in
                ("inserts", "searches", "chain_steps", "compare_bytes", "literals", "matches", "match_bytes", "blocks")]


def build():
    if os.environ.get("ZSO_LIB"):
        return LIB
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("zso_deflate.c", "zso_inflate.c", "zso_adler32.c", "zso_bench.c", "zso.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB


class Oracle:
    def __init__(self, lib):
        self.lib = lib

    def bound(self, n, seg):
        return self.lib.zso_deflate_bound(n, seg)

    def deflate_oneshot(self, data, level, strategy=0, hv=0):
        cap = self.bound(len(data), 0)
        out = ctypes.create_string_buffer(cap)
        n = ctypes.c_size_t(0)
        ad = ctypes.c_uint32(0)
        rc = self.lib.zso_deflate_oneshot(data, len(data), level, strategy, hv, out, cap, ctypes.byref(n), ctypes.byref(ad))
        assert rc == 0, rc
        return out.raw[:n.value], ad.value

    def deflate_segments(self, data, seg_size, level, strategy=0, hv=0):
        """-> (stream bytes, [segment end offsets], adler)"""
        nseg = (len(data) + seg_size - 1) // seg_size
        cap = self.bound(len(data), seg_size)
        out = ctypes.create_string_buffer(cap)
        n = ctypes.c_size_t(0)
        ad = ctypes.c_uint32(0)
        ends = (ctypes.c_uint64 * max(nseg, 1))()
        rc = self.lib.zso_deflate_segments(data, len(data), seg_size, level, strategy, hv, out, cap, ctypes.byref(n),
                                           ends, ctypes.byref(ad))
        assert rc == 0, rc
        return out.raw[:n.value], list(ends[:nseg]), ad.value

    def deflate_writes(self, data, write_size, level, strategy=0, hv=0, flush_mode=3):
        """One Write per write_size bytes with Options.FlushMode = flush_mode (0 none, 1 partial, 2 sync, 3 full), then
        Dispose.  -> (stream bytes, [TotalOut after each Write])"""
        d = self.lib.zso_deflater_new(level, strategy, flush_mode, hv)
        assert d
        ends = []
        for off in range(0, len(data), write_size):
            chunk = data[off:off + write_size]
            assert self.lib.zso_deflater_write(d, chunk, len(chunk)) == 0
            ends.append(self.lib.zso_deflater_total_out(d))
        assert self.lib.zso_deflater_finish(d) == 0
        n = ctypes.c_size_t(0)
        p = self.lib.zso_deflater_output(d, ctypes.byref(n))
        out = ctypes.string_at(p, n.value)
        self.lib.zso_deflater_free(d)
        return out, ends

    def deflate_segments_parallel(self, data, seg_size, level, strategy=0, hv=0, threads=None):
        """deflate_segments on several host threads (ctypes releases the GIL): the stream is cut into contiguous ranges
        of segments, every range after the first is compressed with its PRECEDING segment in front (so the window, the
        Rle halo byte and "not the stream's first segment" are what the single stream has) and that
        segment's output is dropped.  tests/test_oracle_golden.py checks it equals the one-thread call."""
        import zlib
        from concurrent.futures import ThreadPoolExecutor
        n = len(data)
        nseg = (n + seg_size - 1) // seg_size
        threads = threads or min(32, os.cpu_count() or 1)
        per = max(4, (nseg + threads - 1) // threads)
        cuts = list(range(0, nseg, per)) + [nseg]
        if nseg <= per or len(cuts) < 3:
            return self.deflate_segments(data, seg_size, level, strategy, hv)

        def work(ab):
            a, b = ab
            lo = (a - 1) * seg_size if a else 0
            out, ends, _ = self.deflate_segments(data[lo:min(b * seg_size, n)], seg_size, level, strategy, hv)
            if a == 0:
                return out[:ends[-1]], ends
            return out[ends[0]:ends[-1]], [e - ends[0] for e in ends[1:]]

        with ThreadPoolExecutor(max_workers=threads) as ex:
            parts = list(ex.map(work, zip(cuts[:-1], cuts[1:])))
        stream, all_ends, base = [], [], 0
        for piece, ends in parts:
            stream.append(piece)
            all_ends += [base + e for e in ends]
            base += len(piece)
        adler = zlib.adler32(data) & 0xFFFFFFFF
        return b"".join(stream) + b"\x03\x00" + adler.to_bytes(4, "big"), all_ends, adler

    def trace_segments(self, data, seg_size, level, strategy=0, hv=0):
        """Runs the segmented compressor with the block trace on.
        -> (stream, ends, per-segment list of blocks: dict(type, n, stored_len, bit_start, sym=[(dist, lc)...]))"""
        d = self.lib.zso_deflater_new(level, strategy, 3, hv)
        assert d
        segs, cur = [], []

        def cb(_u, blk, dist, lc):
            b = blk.contents
            cur.append(dict(type=b.type, n=b.n_symbols, stored_len=b.stored_len, bit_start=b.bit_start,
                            sym=[(dist[i], lc[i]) for i in range(b.n_symbols)]))

        cbf = BLOCK_CB(cb)
        self.lib.zso_deflater_set_trace(d, cbf, None)
        ends = []
        for off in range(0, len(data), seg_size):
            chunk = data[off:off + seg_size]
            assert self.lib.zso_deflater_write(d, chunk, len(chunk)) == 0
            ends.append(self.lib.zso_deflater_total_out(d))
            segs.append(cur)
            cur = []
        assert self.lib.zso_deflater_finish(d) == 0
        n = ctypes.c_size_t(0)
        p = self.lib.zso_deflater_output(d, ctypes.byref(n))
        out = ctypes.string_at(p, n.value)
        self.lib.zso_deflater_free(d)
        return out, ends, segs

    def counters(self, data, seg_size, level, strategy=0, hv=0):
        d = self.lib.zso_deflater_new(level, strategy, 3 if seg_size else 0, hv)
        step = seg_size or max(len(data), 1)
        for off in range(0, len(data), step):
            chunk = data[off:off + step]
            assert self.lib.zso_deflater_write(d, chunk, len(chunk)) == 0
        assert self.lib.zso_deflater_finish(d) == 0
        c = Counters()
        self.lib.zso_deflater_counters(d, ctypes.byref(c))
        tot = self.lib.zso_deflater_total_out(d)
        self.lib.zso_deflater_free(d)
        return {n: getattr(c, n) for n, _ in Counters._fields_}, tot

    def inflate(self, comp, out_cap):
        out = ctypes.create_string_buffer(max(out_cap, 1))
        n = ctypes.c_size_t(0)
        used = ctypes.c_size_t(0)
        ad = ctypes.c_uint32(0)
        msg = ctypes.c_char_p()
        rc = self.lib.zso_inflate(comp, len(comp), out, out_cap, ctypes.byref(n), ctypes.byref(used), ctypes.byref(ad),
                                  ctypes.byref(msg))
        return rc, out.raw[:n.value], ad.value, (msg.value or b"").decode(), used.value

    def bench_deflate_mt(self, data, seg_size, level, strategy=0, hv=0, threads=1, reps=1):
        """-> (seconds, compressed bytes of one pass): native pinned threads, one contiguous segment range per thread."""
        sec = ctypes.c_double(0)
        ob = ctypes.c_uint64(0)
        rc = self.lib.zso_bench_deflate_mt(data, len(data), seg_size, level, strategy, hv, threads, reps, ctypes.byref(sec), ctypes.byref(ob))
        assert rc == 0, rc
        return sec.value, ob.value

    def bench_inflate_mt(self, comp, seg_ends, first_offset, seg_size, which=1, threads=1, reps=1):
        """-> (seconds, decoded bytes of one pass); which: 1 system zlib, 2 the oracle's inflater."""
        ends = (ctypes.c_uint64 * len(seg_ends))(*seg_ends)
        sec = ctypes.c_double(0)
        dec = ctypes.c_uint64(0)
        rc = self.lib.zso_bench_inflate_mt(comp, ends, len(seg_ends), first_offset, seg_size, which, threads, reps,
                                           ctypes.byref(sec), ctypes.byref(dec))
        assert rc == 0, rc
        return sec.value, dec.value

    def adler32(self, data, seed=1):
        return self.lib.zso_adler32(seed, data, len(data))

    def adler32_combine(self, a1, a2, len2):
        return self.lib.zso_adler32_combine(a1, a2, len2)


def load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB)
        c = ctypes
        lib.zso_deflate_bound.restype = c.c_size_t
        lib.zso_deflate_bound.argtypes = [c.c_size_t, c.c_size_t]
        lib.zso_deflate_oneshot.argtypes = [c.c_char_p, c.c_size_t, c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_size_t,
                                            c.POINTER(c.c_size_t), c.POINTER(c.c_uint32)]
        lib.zso_deflate_segments.argtypes = [c.c_char_p, c.c_size_t, c.c_size_t, c.c_int, c.c_int, c.c_int, c.c_void_p,
                                             c.c_size_t, c.POINTER(c.c_size_t), c.POINTER(c.c_uint64),
                                             c.POINTER(c.c_uint32)]
        lib.zso_deflater_new.restype = c.c_void_p
        lib.zso_deflater_new.argtypes = [c.c_int] * 4
        lib.zso_deflater_free.argtypes = [c.c_void_p]
        lib.zso_deflater_set_trace.argtypes = [c.c_void_p, BLOCK_CB, c.c_void_p]
        lib.zso_deflater_write.argtypes = [c.c_void_p, c.c_char_p, c.c_size_t]
        lib.zso_deflater_finish.argtypes = [c.c_void_p]
        lib.zso_deflater_output.restype = c.c_void_p
        lib.zso_deflater_output.argtypes = [c.c_void_p, c.POINTER(c.c_size_t)]
        lib.zso_deflater_total_out.restype = c.c_uint64
        lib.zso_deflater_total_out.argtypes = [c.c_void_p]
        lib.zso_deflater_total_in.restype = c.c_uint64
        lib.zso_deflater_total_in.argtypes = [c.c_void_p]
        lib.zso_deflater_adler.restype = c.c_uint32
        lib.zso_deflater_adler.argtypes = [c.c_void_p]
        lib.zso_deflater_counters.argtypes = [c.c_void_p, c.POINTER(Counters)]
        lib.zso_inflate.argtypes = [c.c_char_p, c.c_size_t, c.c_void_p, c.c_size_t, c.POINTER(c.c_size_t),
                                    c.POINTER(c.c_size_t), c.POINTER(c.c_uint32), c.POINTER(c.c_char_p)]
        lib.zso_bench_deflate_mt.argtypes = [c.c_char_p, c.c_size_t, c.c_size_t, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int,
                                             c.POINTER(c.c_double), c.POINTER(c.c_uint64)]
        lib.zso_bench_inflate_mt.argtypes = [c.c_char_p, c.POINTER(c.c_uint64), c.c_size_t, c.c_size_t, c.c_size_t, c.c_int, c.c_int,
                                             c.c_int, c.POINTER(c.c_double), c.POINTER(c.c_uint64)]
        lib.zso_adler32.restype = c.c_uint32
        lib.zso_adler32.argtypes = [c.c_uint32, c.c_char_p, c.c_size_t]
        lib.zso_adler32_combine.restype = c.c_uint32
        lib.zso_adler32_combine.argtypes = [c.c_uint32, c.c_uint32, c.c_uint64]
        _lib = Oracle(lib)
    return _lib
