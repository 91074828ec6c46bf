"""Host emulation of the lane-per-segment inflate kernel: the kernel's own __host__ __device__ state machine and table
build (zlibstream_b200/csrc/zsb_inflate_lanes.cuh) compiled for the CPU by tools/inflate_lanes_host and run against the
oracle's streams.  Checks the kernel LOGIC without a GPU; the GPU parity tests proper are in test_gpu_inflate_adler.py."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

import corpus_data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC_DIR = os.path.join(ROOT, "tools", "inflate_lanes_host")
LIB = os.path.join(SRC_DIR, "libzsb_lanes_host.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.fixture(scope="module")
def lanes_host():
    src = os.path.join(SRC_DIR, "inflate_lanes_host.cu")
    deps = [src] + [os.path.join(ROOT, "zlibstream_b200", "csrc", f) for f in
                    ("zsb_inflate_lanes.cuh", "zsb_kernels.cuh", "zsb_common.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        if not os.path.exists(NVCC):
            pytest.skip("nvcc not available")
        subprocess.check_call([NVCC, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                               "-o", LIB, src])
    lib = ctypes.CDLL(LIB)
    lib.zsb_lanes_host_inflate.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint32,
                                           ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t,
                                           ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_int)]

    def inflate(comp, ends, ss, n, misalign=0):
        nseg = len(ends)
        e = (ctypes.c_uint64 * max(nseg, 1))(*ends)
        out = ctypes.create_string_buffer(n + 64)
        lens = (ctypes.c_uint32 * max(nseg, 1))()
        st = (ctypes.c_int * max(nseg, 1))()
        cbuf = ctypes.create_string_buffer(comp + b"\0" * 16, len(comp) + 16)
        lib.zsb_lanes_host_inflate(cbuf, len(comp), e, nseg, 2, ss, ctypes.addressof(out) + misalign, n, lens, st)
        return out.raw[misalign:misalign + n], list(lens), list(st)

    return inflate


def _cases(corpus):
    rng = np.random.RandomState(3)
    return [("corpus", b"".join(corpus[k] for k in sorted(corpus))),
            ("text", corpus_data.text_like(5 * 65536 + 12345, seed=7)),
            ("mixed", corpus_data.mixed_entropy(3 * 262144 + 77, seed=2)),
            ("random", rng.bytes(300001)), ("zeros", bytes(400000)), ("one", b"a"),
            ("sparse", corpus_data.sparse_image(256, 300)), ("png", corpus_data.png_like(300000, seed=3))]


@pytest.mark.parametrize("level,strategy,ss", [(1, 0, 65536), (6, 0, 65536), (9, 0, 262144), (6, 1, 65536), (6, 2, 65536),
                                               (1, 4, 65536), (6, 3, 65536), (3, 0, 131072)])
def test_lane_state_machine_decodes_oracle_streams(lanes_host, oracle, corpus, level, strategy, ss):
    """Every strategy incl. Rle (Q14: a segment may open with a run of the previous segment's last byte), aligned and
    unaligned output."""
    for name, data in _cases(corpus):
        comp, ends, _ = oracle.deflate_segments(data, ss, level, strategy)
        for mis in (0, 3):
            out, lens, st = lanes_host(comp, ends, ss, len(data), mis)
            assert all(s == 0 for s in st), (name, mis, [s for s in st if s][:4])
            assert out == data, (name, mis)
            assert sum(lens) == len(data)


def test_lane_state_machine_survives_corruption(lanes_host, oracle):
    """Bit flips must end in an error status or a (checksum-detectable) wrong output, never a hang or a wild write."""
    rnd = random.Random(5)
    data = corpus_data.text_like(2 * 65536 + 100, seed=9) + bytes(70000) + np.random.RandomState(1).bytes(70000)
    for level, strategy in ((1, 0), (6, 3), (1, 4)):
        comp, ends, _ = oracle.deflate_segments(data, 65536, level, strategy)
        detected = 0
        for _ in range(60):
            bad = bytearray(comp)
            for _ in range(rnd.randint(1, 4)):
                bad[rnd.randrange(2, len(bad) - 6)] ^= 1 << rnd.randrange(8)
            out, lens, st = lanes_host(bytes(bad), ends, 65536, len(data))
            assert all(l <= 65536 for l in lens)
            detected += any(st)
        assert detected > 0
