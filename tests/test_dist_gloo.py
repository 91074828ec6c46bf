"""World-size-2 test of the multi-GPU exchange step (zlibstream_b200/dist.py) on CPU with the gloo backend.

The data path on a GPU box is libzsb200; here a CPU stand-in with the same `deflate_segments_device` signature (built on
the oracle -- test infrastructure) produces each rank's shard, so that what is under test is the host logic the north star
describes: contiguous segment sharding, all_gather of (compressed bytes, input bytes, Adler-32), exclusive scan of the
offsets, variable-size gather to rank 0, adler32_combine fold and trailer.  Rank 0's gathered stream must equal the
oracle's single-process stream for the whole input, byte for byte.
"""
import ctypes
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

SEG = 65536


class OracleShardEngine:
    """CPU stand-in for zlibstream_b200.Engine.deflate_segments_device (host pointers instead of device pointers).

    A full-flush segment k >= 1 compresses to the same bytes whatever preceded it, so a shard that does not
    start the stream is produced by compressing [one filler segment] + shard and dropping the filler's bytes."""

    def __init__(self):
        import oracle_lib
        self.orc = oracle_lib.load()

    def deflate_segments_device(self, in_ptr, n, seg_size, level, strategy, hash_variant, first_segment_index, out_ptr,
                                out_cap, seg_end_ptr=0, header=True, trailer=True, halo_ptr=0):
        data = ctypes.string_at(in_ptr, n)
        if first_segment_index == 0:
            stream, ends, _ = self.orc.deflate_segments(data, seg_size, level, strategy, hash_variant)
            begin, end = (0 if header else 2), (ends[-1] if ends else 2)
        else:
            filler = bytes(seg_size)
            stream, ends, _ = self.orc.deflate_segments(filler + data, seg_size, level, strategy, hash_variant)
            begin, end = ends[0], ends[-1]
            ends = ends[1:]
        body = stream[begin:end] + (stream[end:end + 6] if trailer else b"")
        assert len(body) <= out_cap
        ctypes.memmove(out_ptr, body, len(body))
        if seg_end_ptr:
            arr = (ctypes.c_uint64 * len(ends)).from_address(seg_end_ptr)
            for i, e in enumerate(ends):
                arr[i] = e - begin
        return len(body), self.orc.adler32(data)


class ZlibShardInflater:
    """CPU stand-in for Engine.inflate_segments_device / last_inflate_carry / adler32_device (host pointers).  Stock zlib
    decodes the shard as raw deflate twice, with a one-byte preset dictionary of 0x00 and of 0xFF: the output bytes that
    differ are exactly the Q14 carried run (copies of the byte before the shard), which the device engine leaves unset."""

    def __init__(self):
        self.carry = 0

    def inflate_segments_device(self, in_ptr, n, seg_end_ptr, nseg, seg_size, out_ptr, out_cap, header=True, trailer=True,
                                continuation=False):
        import zlib
        self.out_ptr = out_ptr
        comp = ctypes.string_at(in_ptr, n)
        ends = list((ctypes.c_uint64 * nseg).from_address(seg_end_ptr)) if nseg else []
        body = comp[2 if header else 0:ends[-1] if ends else (2 if header else 0)]
        outs = []
        for fill in (b"\x00", b"\xff"):
            d = zlib.decompressobj(wbits=-15, zdict=fill)
            outs.append(d.decompress(body))
        a, b = outs
        assert len(a) == len(b) <= out_cap
        self.carried = [i for i in range(len(a)) if a[i] != b[i]]   # the first segment's prefix, and whole segments after it
        self.carry = next((i for i in range(len(a)) if a[i] == b[i]), len(a)) if a != b else 0
        assert continuation or not self.carried
        self.n_out = len(a)
        ctypes.memmove(out_ptr, a, len(a))
        return len(a), zlib.adler32(a) & 0xFFFFFFFF

    def last_inflate_carry(self):
        return self.carry

    def inflate_fill_carry(self, prev_byte):
        import zlib
        buf = (ctypes.c_uint8 * self.n_out).from_address(self.out_ptr)
        for i in self.carried:
            buf[i] = prev_byte
        return zlib.adler32(bytes(buf)) & 0xFFFFFFFF


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, level, n, gather, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import corpus_data
        import oracle_lib
        from zlibstream_b200.dist import deflate_sharded, shard_segments
        orc = oracle_lib.load()
        data = corpus_data.text_like(n, 3)
        nseg = (n + SEG - 1) // SEG
        b, e = shard_segments(nseg, world, rank)
        shard = data[b * SEG:min(e * SEG, n)]
        local_in = torch.frombuffer(bytearray(shard), dtype=torch.uint8)
        cap = orc.bound(len(shard), SEG)
        local_out = torch.zeros(cap, dtype=torch.uint8)
        seg_end = torch.zeros(max(e - b, 1), dtype=torch.int64)
        want, want_ends, want_adler = orc.deflate_segments(data, SEG, level)
        gather_to = None
        if gather:
            gather_to = torch.zeros(len(want) + 64, dtype=torch.uint8) if rank == 0 else torch.zeros(1, dtype=torch.uint8)
        res = deflate_sharded(OracleShardEngine(), orc.adler32_combine, local_in, SEG, level, 0, 0, b, local_out, seg_end,
                              gather_to=gather_to)
        assert res.total_len == len(want), (res.total_len, len(want))
        assert res.adler == want_adler
        assert len(res.sizes) == world and res.offsets[0] == 0
        # every rank's offset is the reference's TotalOut at its first segment
        my_first = want_ends[b - 1] if b else 0
        assert res.offsets[rank] == my_first, (rank, res.offsets, my_first)
        if rank == 0 and gather:
            got = bytes(res.stream.numpy().tobytes())
            assert got == want
        with open(result_path + f".{rank}", "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def _inflate_worker(rank, world, port, kind, gather, bad_adler, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import zlib
        import corpus_data
        import oracle_lib
        from zlibstream_b200.dist import inflate_sharded, shard_index
        from zlibstream_b200.options import ZlibStreamException
        orc = oracle_lib.load()
        if kind == "text":
            data, level, strategy = corpus_data.text_like(5 * SEG + 999, 5), 6, 0
        else:   # Rle: runs that cross every segment boundary (Q14), one shard that is a single run from end to end
            data, level, strategy = b"A" * (SEG + 100) + b"B" * (3 * SEG) + corpus_data.text_like(SEG, 2), 1, 3
        stream, ends, adler = orc.deflate_segments(data, SEG, level, strategy)
        b, e, lo, hi = shard_index(ends, 2, world, rank)
        comp = torch.frombuffer(bytearray(stream[lo:hi]), dtype=torch.uint8)
        local_ends = torch.tensor([x - lo for x in ends[b:e]], dtype=torch.int64)
        local_out = torch.zeros((e - b) * SEG, dtype=torch.uint8)
        gather_to = None
        if gather:
            gather_to = torch.zeros(len(data) + 16, dtype=torch.uint8) if rank == 0 else torch.zeros(1, dtype=torch.uint8)
        want = adler ^ 1 if bad_adler else adler
        try:
            res = inflate_sharded(ZlibShardInflater(), orc.adler32_combine, comp, local_ends, SEG, local_out, has_header=(b == 0),
                                  expect_adler=want, gather_to=gather_to)
            assert not bad_adler
        except ZlibStreamException as ex:
            assert bad_adler and "incorrect data check" in str(ex)
            res = None
        if res is not None:
            assert res.adler == zlib.adler32(data) & 0xFFFFFFFF and res.total_len == len(data)
            assert res.offsets[rank] == b * SEG
            mine = data[b * SEG:min(e * SEG, len(data))]
            assert bytes(local_out[:res.local_len].numpy().tobytes()) == mine
            if kind == "rle" and rank > 0:
                assert res.carry > 0     # the shard really opened with a carried run
            if rank == 0 and gather:
                assert bytes(res.out.numpy().tobytes()) == data
        with open(result_path + f".{rank}", "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,gather,bad_adler", [(2, "text", True, False), (2, "rle", True, False), (3, "rle", False, False),
                                                        (2, "text", False, True)])
def test_sharded_inflate_matches_the_input(tmp_path, world, kind, gather, bad_adler):
    """inflate shards by compressed-offset ranges; Adler-32 folds with adler32_combine; Q14 carries cross ranks."""
    port = _free_port()
    result = str(tmp_path / "res")
    mp.spawn(_inflate_worker, args=(world, port, kind, gather, bad_adler, result), nprocs=world, join=True)
    for r in range(world):
        assert os.path.exists(result + f".{r}")


@pytest.mark.parametrize("level,n,gather", [(1, 5 * SEG + 1234, True), (6, 4 * SEG, True), (1, 3 * SEG + 7, False)])
def test_two_rank_sharded_deflate_matches_single_stream(tmp_path, level, n, gather):
    world = 2
    port = _free_port()
    result = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, port, level, n, gather, result), nprocs=world, join=True)
    for r in range(world):
        assert os.path.exists(result + f".{r}")


def test_shard_segments_covers_everything():
    from zlibstream_b200.dist import exclusive_scan, shard_segments
    for nseg in (0, 1, 2, 7, 8, 9, 16384):
        for world in (1, 2, 3, 8):
            ranges = [shard_segments(nseg, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nseg
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in ranges) - min(e - b for b, e in ranges) <= 1
    assert exclusive_scan([3, 4, 5]) == [0, 3, 7]
