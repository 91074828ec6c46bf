"""ctypes binding of libzsb200.so (include/zsb200.h).  There is no fallback: a missing library is an error."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZSB200_LIB") or os.path.join(_HERE, "libzsb200.so")   # ZSB200_LIB: A/B builds for profiling

ZSB_IN_DEVICE = 0x1
ZSB_OUT_DEVICE = 0x2
ZSB_NO_HEADER = 0x4
ZSB_NO_TRAILER = 0x8
ZSB_OFFSETS_DEVICE = 0x10
ZSB_CONTINUATION = 0x20
ZSB_SINGLE_STREAM = 0x40
ZSB_SYNC_FLUSH = 0x80
ZSB_PARTIAL_FLUSH = 0x100

# every symbol include/zsb200.h declares
EXPORTS = [
    "zsb_ctx_create", "zsb_ctx_destroy", "zsb_last_error", "zsb_ctx_set_stream", "zsb_ctx_get_stream",
    "zsb_ctx_launch_count", "zsb_ctx_last_kernel_ms", "zsb_ctx_last_kernel_ms_detail", "zsb_deflate_bound", "zsb_deflate_segments",
    "zsb_inflate_segments", "zsb_ctx_last_inflate_carry", "zsb_inflate_fill_carry", "zsb_build_segment_index", "zsb_png_bound", "zsb_png_encode", "zsb_adler32", "zsb_adler32_combine", "zsb_version", "zsb_debug_fetch",
]

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    """Loads libzsb200.so and declares prototypes.  Raises NativeLibraryMissing if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(zlibstream_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, sz, u64, u32, i32 = c.c_void_p, c.c_size_t, c.c_uint64, c.c_uint32, c.c_int
    lib.zsb_ctx_create.argtypes = [i32, c.POINTER(vp)]
    lib.zsb_ctx_create.restype = i32
    lib.zsb_ctx_destroy.argtypes = [vp]
    lib.zsb_ctx_destroy.restype = None
    lib.zsb_last_error.argtypes = [vp]
    lib.zsb_last_error.restype = c.c_char_p
    lib.zsb_ctx_set_stream.argtypes = [vp, vp]
    lib.zsb_ctx_set_stream.restype = i32
    lib.zsb_ctx_get_stream.argtypes = [vp]
    lib.zsb_ctx_get_stream.restype = vp
    lib.zsb_ctx_launch_count.argtypes = [vp]
    lib.zsb_ctx_launch_count.restype = u64
    lib.zsb_ctx_last_kernel_ms.argtypes = [vp, c.POINTER(c.c_float), c.POINTER(c.c_float), c.POINTER(c.c_float)]
    lib.zsb_ctx_last_kernel_ms.restype = i32
    lib.zsb_ctx_last_kernel_ms_detail.argtypes = [vp, c.POINTER(c.c_float), i32]
    lib.zsb_ctx_last_kernel_ms_detail.restype = i32
    lib.zsb_deflate_bound.argtypes = [sz, sz]
    lib.zsb_deflate_bound.restype = sz
    lib.zsb_deflate_segments.argtypes = [vp, vp, sz, sz, i32, i32, i32, u64, vp, c.c_uint, vp, sz, c.POINTER(sz), vp,
                                         c.POINTER(u32)]
    lib.zsb_deflate_segments.restype = i32
    lib.zsb_inflate_segments.argtypes = [vp, vp, sz, vp, sz, sz, c.c_uint, vp, sz, c.POINTER(sz), c.POINTER(u32)]
    lib.zsb_inflate_segments.restype = i32
    lib.zsb_ctx_last_inflate_carry.argtypes = [vp, c.POINTER(u32)]
    lib.zsb_ctx_last_inflate_carry.restype = i32
    lib.zsb_inflate_fill_carry.argtypes = [vp, c.c_uint8, c.POINTER(u32)]
    lib.zsb_inflate_fill_carry.restype = i32
    lib.zsb_build_segment_index.argtypes = [vp, vp, sz, sz, c.c_uint, vp, sz, c.POINTER(sz), c.POINTER(u64)]
    lib.zsb_build_segment_index.restype = i32
    lib.zsb_png_bound.argtypes = [u32, u32, i32, sz, sz]
    lib.zsb_png_bound.restype = sz
    lib.zsb_png_encode.argtypes = [vp, vp, u32, u32, i32, i32, i32, i32, sz, sz, c.c_uint, vp, sz, c.POINTER(sz)]
    lib.zsb_png_encode.restype = i32
    lib.zsb_adler32.argtypes = [vp, u32, vp, sz, c.c_uint, c.POINTER(u32)]
    lib.zsb_adler32.restype = i32
    lib.zsb_adler32_combine.argtypes = [u32, u32, u64]
    lib.zsb_adler32_combine.restype = u32
    lib.zsb_version.argtypes = []
    lib.zsb_version.restype = c.c_char_p
    lib.zsb_debug_fetch.argtypes = [vp, i32, u64, vp, sz]
    lib.zsb_debug_fetch.restype = i32
    _lib = lib
    return lib
