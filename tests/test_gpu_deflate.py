"""GPU parity: libzsb200's segmented deflate must be bit-identical to the oracle (= the reference's dialect) for the
same level / strategy / hash variant / full-flush segment boundaries.  All calls go through the C ABI."""
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

import corpus_data

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SEGMENTED = json.load(open(os.path.join(HERE, "golden", "segmented.json")))


def _blob(corpus):
    return b"".join(corpus[k] for k in sorted(corpus))


@pytest.mark.parametrize("key", sorted(SEGMENTED))
def test_golden_segmented_vectors(engine, corpus, key):
    """Committed oracle outputs (tests/golden/segmented.json): L1/3/6/9, Filtered, HuffmanOnly, Rle, Fixed; 64 / 256 KiB."""
    _, seg, lvl, strat = key.split("/")
    ss, level, strategy = int(seg[3:]), int(lvl[1:]), int(strat[1:])
    data = _blob(corpus)
    got, ends, adler = engine.deflate_segments(data, ss, level, strategy)
    g = SEGMENTED[key]
    assert len(got) == g["len"] and hashlib.sha256(got).hexdigest() == g["sha256"]
    assert adler == g["adler"] == zlib.adler32(data)
    assert hashlib.sha256(",".join(map(str, ends)).encode()).hexdigest() == g["ends_sha256"]


@pytest.mark.parametrize("n", [65536, 65536 + 32505, 65536 + 32506, 3 * 65536 + 32767, 5 * 65536, 2 * 262144 + 40000, 131072 + 1])
@pytest.mark.parametrize("ss", [65536, 262144])
def test_level0_stored_blocks(engine, oracle, n, ss):
    """SURVEY 8(f)-3, DeflateStored (Deflate.Stored.cs:24-84): 32 KiB stored blocks, the empty stored block the end-of-input
    flush leaves behind a multiple of 32 KiB or a tail >= w_size - MIN_LOOKAHEAD, empty STORED blocks where Q10 repeats."""
    data = corpus_data.text_like(n, seed=n % 97)
    want, want_ends, want_adler = oracle.deflate_segments(data, ss, 0, 0)
    got, ends, adler = engine.deflate_segments(data, ss, 0, 0)
    assert got == want and ends == want_ends and adler == want_adler
    assert zlib.decompress(got) == data
    dec, ad = engine.inflate_segments(got, ends, ss, len(data))
    assert dec == data and ad == adler
    # one Write, NoFlush, Dispose at level 0
    one, _ = engine.deflate_stream(data, 0)
    assert one == oracle.deflate_oneshot(data, 0)[0]
    # Filtered / HuffmanOnly / Fixed do not change level 0 (Rle would override the level's function, Q13: not built at level 0)
    for strategy in (1, 2, 4):
        assert engine.deflate_segments(data, ss, 0, strategy)[0] == oracle.deflate_segments(data, ss, 0, strategy)[0]
    with pytest.raises(Exception, match="not implemented"):
        engine.deflate_segments(data, ss, 0, 3)


@pytest.mark.parametrize("level,strategy", [(1, 0), (3, 0), (6, 0), (9, 0), (6, 1), (6, 2), (6, 3), (1, 4)])
@pytest.mark.parametrize("ss", [65536, 262144])
def test_sync_flush_segments(engine, oracle, corpus, level, strategy, ss):
    """SURVEY 8(f)-3, FlushMode.SyncFlush on every Write (Deflate.cs:583-595): the marker of a full flush, but the hash table
    stays, so matches reach back into the previous segment.  The stream is one piece for an inflater."""
    data = (corpus["alice29.txt"] + corpus["ptt5"] + corpus["alice29.txt"] + corpus["kennedy.xls"])[:5 * 262144 + 777]
    want, want_ends = oracle.deflate_writes(data, ss, level, strategy, 0, flush_mode=2)
    full, _, _ = oracle.deflate_segments(data, ss, level, strategy)
    got, ends, adler = engine.deflate_segments(data, ss, level, strategy, sync_flush=True)
    assert got == want and ends == want_ends and adler == zlib.adler32(data)
    if strategy in (0, 1, 4):
        assert len(want) < len(full)          # the second copy of alice29 really is found across the flush points
    assert zlib.decompress(got) == data
    dec, ad = engine.inflate_stream(got, len(data))
    assert dec == data and ad == adler


@pytest.mark.parametrize("level,strategy", [(1, 0), (3, 0), (6, 0), (9, 0), (6, 1), (6, 2), (6, 3), (1, 4)])
@pytest.mark.parametrize("ss", [65536, 262144])
def test_partial_flush_segments(engine, oracle, corpus, level, strategy, ss):
    """SURVEY 8(f)-3, FlushMode.PartialFlush on every Write (Deflate.cs:566-581, Tr_align Trees.cs:658-680): one or two empty
    static blocks instead of the stored marker, the stream not byte aligned between Writes, TotalOut = whole bytes so far."""
    data = (corpus["alice29.txt"] + corpus["ptt5"] + corpus["alice29.txt"] + corpus["kennedy.xls"])[:5 * 262144 + 777]
    want, want_ends = oracle.deflate_writes(data, ss, level, strategy, 0, flush_mode=1)
    got, ends, adler = engine.deflate_segments(data, ss, level, strategy, partial_flush=True)
    assert got == want and ends == want_ends and adler == zlib.adler32(data)
    assert zlib.decompress(got) == data
    dec, ad = engine.inflate_stream(got, len(data))
    assert dec == data and ad == adler


@pytest.mark.parametrize("level", [1, 2, 3, 4, 5, 6, 7, 8, 9, -1])
def test_every_level_matches_oracle(engine, oracle, corpus, level):
    data = corpus["lcet10.txt"] + corpus["sum"] + corpus["ptt5"][:200000]
    want, want_ends, want_adler = oracle.deflate_segments(data, 65536, level)
    got, ends, adler = engine.deflate_segments(data, 65536, level)
    assert got == want and ends == want_ends and adler == want_adler


@pytest.mark.parametrize("strategy", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("level", [1, 6])
def test_every_strategy_matches_oracle(engine, oracle, corpus, level, strategy):
    data = corpus["kennedy.xls"][:400000] + corpus_data.png_like(300000) + corpus["alice29.txt"]
    want, want_ends, _ = oracle.deflate_segments(data, 65536, level, strategy)
    got, ends, _ = engine.deflate_segments(data, 65536, level, strategy)
    assert got == want and ends == want_ends


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 11, 261, 262, 263, 32767, 32768, 32769, 65535, 65536, 65537, 65536 + 32768,
                               131072, 131073, 3 * 65536 + 258])
def test_ragged_and_tiny_inputs(engine, oracle, corpus, n):
    """Empty input, sub-window inputs, inputs ending right at / next to refill and segment boundaries."""
    data = (corpus["asyoulik.txt"] + corpus["cp.html"])[:n]
    for level in (1, 6):
        want, want_ends, _ = oracle.deflate_segments(data, 65536, level)
        got, ends, _ = engine.deflate_segments(data, 65536, level)
        assert got == want and ends == want_ends, (n, level)


@pytest.mark.parametrize("kind", ["zeros", "random", "runs", "sparse"])
def test_degenerate_data(engine, oracle, kind):
    rng = np.random.RandomState(5)
    n = 5 * 65536 + 1000
    if kind == "zeros":
        data = bytes(n)                       # 258-byte matches everywhere, Q3 two-cycle in prev[]
    elif kind == "random":
        data = rng.bytes(n)                   # stored blocks
    elif kind == "runs":
        data = bytes(np.repeat(rng.randint(0, 256, size=n // 40 + 1).astype(np.uint8), 40)[:n])
    else:
        data = corpus_data.sparse_image(512, 700)[:n]
    for level, strategy in ((1, 0), (6, 0), (9, 0), (6, 3), (6, 2), (6, 1)):
        want, want_ends, _ = oracle.deflate_segments(data, 65536, level, strategy)
        got, ends, _ = engine.deflate_segments(data, 65536, level, strategy)
        assert got == want and ends == want_ends, (kind, level, strategy)


def test_symbol_count_hits_block_limit_exactly(engine, oracle):
    """HuffmanOnly: every full block is exactly 16383 literals (Q9); 4 * 16383 + 4 = 65536 exercises the boundary."""
    rng = np.random.RandomState(9)
    data = bytes(rng.randint(32, 127, size=4 * 65536, dtype=np.uint8))
    for ss in (65536, 131072):
        want, want_ends, _ = oracle.deflate_segments(data, ss, 6, 2)
        got, ends, _ = engine.deflate_segments(data, ss, 6, 2)
        assert got == want and ends == want_ends


def test_256k_segments_levels_6_and_9(engine, oracle):
    data = corpus_data.mixed_entropy(6 * 262144 + 5000)     # BASELINE config 3 shape
    for level in (6, 9):
        want, want_ends, _ = oracle.deflate_segments(data, 262144, level)
        got, ends, _ = engine.deflate_segments(data, 262144, level)
        assert got == want and ends == want_ends


def test_multiplicative_hash_variant(engine, oracle, corpus):
    data = corpus["plrabn12.txt"][:300000]
    for level in (1, 6):
        want, _, _ = oracle.deflate_segments(data, 65536, level, 0, 1)
        got, _, _ = engine.deflate_segments(data, 65536, level, 0, 1)
        assert got == want


def test_sharded_ranges_concatenate_to_the_single_stream(engine, oracle, corpus):
    """What the multi-GPU path relies on: compressing contiguous segment ranges separately (first_segment_index, no
    header / trailer on later shards) and concatenating gives the single-stream bytes."""
    data = _blob(corpus)
    ss = 65536
    want, want_ends, want_adler = oracle.deflate_segments(data, ss, 6)
    nseg = (len(data) + ss - 1) // ss
    cuts = [0, 7, 8, 30, nseg]
    parts, adler = [], 1
    from zlibstream_b200 import adler32_combine
    for a, b in zip(cuts[:-1], cuts[1:]):
        chunk = data[a * ss:b * ss]
        halo = data[a * ss - 32768:a * ss] if a else None
        got, ends, ad = engine.deflate_segments(chunk, ss, 6, 0, first_segment_index=a, halo=halo, header=(a == 0), trailer=False)
        parts.append(got)
        adler = adler32_combine(adler, ad, len(chunk))
    stream = b"".join(parts) + b"\x03\x00" + adler.to_bytes(4, "big")
    assert stream == want and adler == want_adler


def test_rle_needs_the_halo_byte(engine, oracle, corpus):
    """Q14: an Rle segment may start with a distance-1 copy of the previous segment's last byte."""
    data = corpus["ptt5"]
    ss = 65536
    want, _, _ = oracle.deflate_segments(data, ss, 6, 3)
    a = 3
    head, _, ad0 = engine.deflate_segments(data[:a * ss], ss, 6, 3, trailer=False)
    tail, _, ad1 = engine.deflate_segments(data[a * ss:], ss, 6, 3, first_segment_index=a, halo=data[a * ss - 32768:a * ss],
                                           header=False, trailer=False)
    assert (head + tail) == want[:-6]


def test_output_buffer_too_small_is_buf_error(engine, corpus):
    import ctypes
    from zlibstream_b200 import _native
    lib = _native.load()
    data = corpus["alice29.txt"]
    out = ctypes.create_string_buffer(1000)
    n = ctypes.c_size_t(0)
    ad = ctypes.c_uint32(0)
    rc = lib.zsb_deflate_segments(engine._h, data, len(data), 65536, 6, 0, 0, 0, None, 0, out, 1000, ctypes.byref(n), None,
                                  ctypes.byref(ad))
    assert rc == -5 and engine.message == "buffer error" and n.value > 1000    # CompressionState.ZBUFERROR


def test_argument_errors(engine):
    from zlibstream_b200 import ZlibStreamException
    for kw in (dict(level=10), dict(level=-2), dict(strategy=5), dict(seg_size=65536 + 512), dict(seg_size=32768)):
        args = dict(seg_size=65536, level=6, strategy=0)
        args.update(kw)
        with pytest.raises(ZlibStreamException, match="deflating: "):
            engine.deflate_segments(b"x" * 100, args["seg_size"], args["level"], args["strategy"])


def test_large_property_round_trip(engine):
    """BASELINE-size shape at reduced length (256 MiB): size-independent properties -- stock zlib decodes the stream to
    the input, every segment ends with the marker, Adler-32 matches, the device inflater round-trips it."""
    import torch
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import synth
    n = 256 << 20
    ss = 65536
    data = synth.text_like(n, seed=11, device="cuda")
    cap = engine.deflate_bound(n, ss)
    out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    ends = torch.empty(n // ss, dtype=torch.int64, device="cuda")
    out_len, adler = engine.deflate_segments_device(data.data_ptr(), n, ss, 1, 0, 0, 0, out.data_ptr(), cap, ends.data_ptr())
    host = data.cpu().numpy().tobytes()
    comp = out[:out_len].cpu().numpy().tobytes()
    assert adler == zlib.adler32(host)
    e = ends.cpu().numpy()
    assert all(comp[int(x) - 4:int(x)] == b"\x00\x00\xff\xff" for x in e[::97])
    assert zlib.decompress(comp) == host
    dec = torch.empty(n, dtype=torch.uint8, device="cuda")
    dl, dad = engine.inflate_segments_device(out.data_ptr(), out_len, ends.data_ptr(), n // ss, ss, dec.data_ptr(), n)
    assert dl == n and dad == adler and torch.equal(dec, data)


def _run_lengths_data(seed, n):
    """Runs of every length around the Rle thresholds (1..3, 257..262, 515..520, long), separated by distinct bytes."""
    rng = np.random.RandomState(seed)
    out = bytearray()
    lens = list(range(1, 8)) + list(range(255, 264)) + list(range(513, 522)) + [1000, 5000, 40000]
    while len(out) < n:
        r = int(lens[rng.randint(len(lens))])
        b = int(rng.randint(256))
        if out and out[-1] == b:
            b = (b + 1) & 255
        out += bytes([b]) * r
        if rng.randint(3) == 0:
            out += rng.bytes(int(rng.randint(1, 40)))
    return bytes(out[:n])


@pytest.mark.parametrize("ss", [65536, 262144])
def test_parallel_symbolizer_rle_and_huffman_only(engine, oracle, ss):
    """zsb_symbolize.cu: Rle run arithmetic (remainders below 3, exact multiples of 258, runs across segment boundaries,
    Q14), HuffmanOnly with both parsers, block cuts at 16383 symbols incl. the empty trailing block, buf_valid."""
    nolit = bytes((i * 7 + (i >> 8)) & 255 for i in range(ss + 2 * 16383))          # no equal neighbours: symbols == bytes
    cases = [_run_lengths_data(1, 3 * ss + 12345), _run_lengths_data(2, ss), bytes(2 * ss + 777), b"\x07" * (ss + 259),
             nolit, nolit[:ss + 16383], np.random.RandomState(5).bytes(ss + 100), b"ab" * 40000 + bytes(3) + b"c" * 70000]
    for data in cases:
        for level, strategy in ((6, 3), (1, 3), (6, 2), (1, 2)):
            want, want_ends, want_adler = oracle.deflate_segments(data, ss, level, strategy)
            got, ends, adler = engine.deflate_segments(data, ss, level, strategy)
            assert got == want and ends == want_ends and adler == want_adler, (len(data), level, strategy)


@pytest.mark.parametrize("level", [4, 5, 7, 8])
def test_slow_pipeline_stress(engine, oracle, corpus, level):
    """Levels 4-9 run as a three-stage pipeline over 992-position tiles with a shadow of reused prev[] slots
    (zsb_lz77.cu): long runs (matches that skip past tile ends and past the pre-inserted region), incompressible
    stretches, short last segments, refills in the middle of a tile, 128 KiB and 256 KiB segments."""
    rng = np.random.RandomState(level)
    blob = b"".join(corpus[k] for k in sorted(corpus))
    parts = [blob[:300000], bytes(70000), rng.bytes(50000), b"abcdefgh" * 9000, blob[300000:500000], b"\x01" * 1000,
             _run_lengths_data(level, 200000), corpus_data.mixed_entropy(400000, seed=level)]
    data = b"".join(parts)
    for ss, strategy in ((65536, 0), (131072, 0), (262144, 1)):
        want, want_ends, want_adler = oracle.deflate_segments(data, ss, level, strategy)
        got, ends, adler = engine.deflate_segments(data, ss, level, strategy)
        assert got == want and ends == want_ends and adler == want_adler, (level, ss, strategy)


def test_small_scratch_batches(engine, oracle, monkeypatch):
    """ZSB_SCRATCH_BYTES bounds the symbol scratch, so a large input goes through the kernels in several batches of
    segments: batch boundaries must not show (running output offsets, Adler fold, the byte before a batch for Rle / Q14,
    the previous window for a short last segment / Q7)."""
    monkeypatch.setenv("ZSB_SCRATCH_BYTES", str(3 * (65536 * 4 + 40000)))   # about three 64 KiB segments per batch
    data = _run_lengths_data(9, 8 * 65536) + b"\x05" * 70000 + corpus_data.text_like(3 * 65536 + 1234, seed=3)
    for level, strategy in ((6, 3), (1, 0), (6, 0), (6, 2), (9, 1)):
        want, want_ends, want_adler = oracle.deflate_segments(data, 65536, level, strategy)
        got, ends, adler = engine.deflate_segments(data, 65536, level, strategy)
        assert got == want and ends == want_ends and adler == want_adler, (level, strategy)


def test_unaligned_device_pointers(engine, oracle):
    """Device-pointer entry points with the input at an odd address and the output 4-byte (not 16-byte) aligned (the
    kernels' aligned fast paths must fall back): deflate of every parser family, then both inflate kernels into an
    odd-address output.  A device output pointer that is not 4-byte aligned is refused."""
    import torch
    data = _run_lengths_data(4, 2 * 65536) + corpus_data.text_like(3 * 65536 + 777, seed=8)
    n, seg = len(data), 65536
    nseg = (n + seg - 1) // seg
    buf = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    buf[3:3 + n] = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    cap = engine.deflate_bound(n, seg)
    out = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
    ends = torch.zeros(nseg, dtype=torch.int64, device="cuda")
    dec = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    for level, strategy in ((1, 0), (6, 0), (6, 2), (6, 3)):
        want, want_ends, _ = oracle.deflate_segments(data, seg, level, strategy)
        clen, _ = engine.deflate_segments_device(buf.data_ptr() + 3, n, seg, level, strategy, 0, 0, out.data_ptr() + 4, cap,
                                                 ends.data_ptr())
        assert bytes(out[4:4 + clen].cpu().numpy()) == want and ends.cpu().tolist() == want_ends, (level, strategy)
        for kernel in ("warp", "lanes"):
            import os
            os.environ["ZSB_INFLATE_KERNEL"] = kernel
            try:
                dec.zero_()
                got_n, _ = engine.inflate_segments_device(out.data_ptr() + 4, clen, ends.data_ptr(), nseg, seg,
                                                          dec.data_ptr() + 1, n)
            finally:
                del os.environ["ZSB_INFLATE_KERNEL"]
            assert got_n == n and bytes(dec[1:1 + n].cpu().numpy()) == data, (level, strategy, kernel)
    from zlibstream_b200 import ZlibStreamException
    with pytest.raises(ZlibStreamException, match="4-byte aligned"):
        engine.deflate_segments_device(buf.data_ptr() + 3, n, seg, 1, 0, 0, 0, out.data_ptr() + 5, cap, ends.data_ptr())


def test_piped_host_path_equals_device_path(engine, corpus):
    """With host buffers and enough segments, zsb_deflate_segments uploads / compresses / downloads in overlapping
    batches.  Its output must be byte for byte what the single-batch device-pointer path produces (which the other tests
    pin to the oracle), for a parser with a window history (level 1), for Rle (Q14 across batches) and a short last segment."""
    import torch
    blob = b"".join(corpus[k] for k in sorted(corpus))
    seg = 65536
    data = (blob * (170 * (1 << 20) // len(blob) + 1))[:2500 * seg + 12345]
    n = len(data)
    nseg = (n + seg - 1) // seg
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    cap = engine.deflate_bound(n, seg)
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_ends = torch.empty(nseg, dtype=torch.int64, device="cuda")
    for level, strategy in ((1, 0), (6, 3)):
        clen, d_adler = engine.deflate_segments_device(d_in.data_ptr(), n, seg, level, strategy, 0, 0, d_out.data_ptr(), cap,
                                                       d_ends.data_ptr())
        want = bytes(d_out[:clen].cpu().numpy())
        got, ends, adler = engine.deflate_segments(data, seg, level, strategy)
        assert len(got) == len(want) and got == want, (level, strategy)
        assert ends == d_ends.cpu().tolist() and adler == d_adler == zlib.adler32(data)


# ---- parity where the numbers are taken: several segments per CTA of the persistent kernels ---------------------------

def _multiwave_input(kind, n):
    """Deterministic inputs in the tens of megabytes, built by tiling + perturbing the small generators (every segment
    differs from the others, so a stale window / hash table / tile buffer of the previous segment would show)."""
    rng = np.random.RandomState(77)
    if kind == "text":
        base = corpus_data.text_like(4 << 20, seed=21)
    elif kind == "mixed":
        base = corpus_data.mixed_entropy(5 * 262144 * 3, seed=22)
    else:
        base = corpus_data.png_like(4 << 20, seed=23)
    a = np.frombuffer(base, dtype=np.uint8)
    reps = (n + a.size - 1) // a.size
    out = np.tile(a, reps)[:n].copy()
    # one random byte changed about every 4 KiB, so that the copies of the base are not identical segments
    pos = rng.randint(0, n, size=n // 4096)
    out[pos] = rng.randint(0, 256, size=pos.size).astype(np.uint8)
    return out.tobytes()


MULTIWAVE = [
    # (name, data kind, segment size, #segments, level, strategy): >= 3 x 148 segments of 64 KiB, >= 2 x 148 of 256 KiB
    ("L1-64K", "text", 65536, 460, 1, 0), ("L3-64K", "text", 65536, 460, 3, 0), ("L6-64K", "text", 65536, 460, 6, 0),
    ("L9-64K", "text", 65536, 460, 9, 0), ("Filtered-64K", "png", 65536, 460, 6, 1), ("Rle-64K", "png", 65536, 460, 6, 3),
    ("HuffmanOnly-64K", "png", 65536, 460, 6, 2), ("L1-256K", "mixed", 262144, 310, 1, 0),
    ("L6-256K", "mixed", 262144, 310, 6, 0), ("L9-256K", "text", 262144, 310, 9, 0), ("Rle-256K", "png", 262144, 310, 6, 3),
    ("L2-64K-mixed", "mixed", 65536, 460, 2, 0),
]


@pytest.mark.parametrize("name,kind,ss,nseg,level,strategy", MULTIWAVE, ids=[m[0] for m in MULTIWAVE])
def test_multiwave_parity_with_the_oracle(engine, oracle, name, kind, ss, nseg, level, strategy):
    """The LZ77 kernels are persistent (grid = #SMs, dynamic work counter): with more than 3 x 148 segments every CTA
    parses several segments back to back, so state carried from one to the next (window, head / prev, slot table,
    tile buffers, pipeline flags -- what a full flush does and does not reset, Deflate.cs:583-614) is compared with the
    oracle bit for bit.  The same streams are then decoded by both inflate kernels."""
    data = _multiwave_input(kind, ss * nseg + 1234)
    want, want_ends, want_adler = oracle.deflate_segments_parallel(data, ss, level, strategy)
    got, ends, adler = engine.deflate_segments(data, ss, level, strategy)
    assert len(got) == len(want) and hashlib.sha256(got).digest() == hashlib.sha256(want).digest()
    assert ends == want_ends and adler == want_adler
    for kernel in ("warp", "lanes"):
        os.environ["ZSB_INFLATE_KERNEL"] = kernel
        try:
            dec, dad = engine.inflate_segments(want, want_ends, ss, len(data))
        finally:
            del os.environ["ZSB_INFLATE_KERNEL"]
        assert dad == want_adler and dec == data, (name, kernel)


@pytest.mark.parametrize("kernel", ["lanes", "lanes_small"])
@pytest.mark.parametrize("per_warp", [1, 3])
def test_lane_inflate_with_several_warps_per_cta(engine, oracle, monkeypatch, kernel, per_warp):
    """The lane kernels run one CTA per SM whose warps share ONE table-build scratch behind a lock
    (zsb_inflate_lanes.cu).  With few segments per warp, 460 segments need more warps than there are SMs, so every CTA
    holds several warps that build their tables at the same time (dynamic Huffman blocks: a build per block and lane)."""
    ss, nseg = 65536, 460
    data = _multiwave_input("text", ss * nseg + 77)
    want, want_ends, want_adler = oracle.deflate_segments_parallel(data, ss, 6, 0)
    monkeypatch.setenv("ZSB_INFLATE_KERNEL", kernel)
    monkeypatch.setenv("ZSB_INFLATE_SEGS_PER_WARP", str(per_warp))
    dec, dad = engine.inflate_segments(want, want_ends, ss, len(data))
    assert dad == want_adler and dec == data
