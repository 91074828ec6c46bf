"""Multi-GPU segmented deflate and inflate: one process per GPU, segments sharded contiguously across ranks.

The reference has no multi-process code at all; this is the exchange step the north star
describes for full-flush segments, which are independent units of work:

  1. every rank compresses its contiguous range of segments with the device engine (no collective on the data path);
     only stream segment 0 (rank 0's first) is "first" (header + 510-byte first chunk);
  2. all_gather of (compressed bytes, input bytes, Adler-32) per rank  -- NCCL over NVLink on GPUs, gloo in CPU tests;
  3. exclusive scan of the compressed sizes -> every rank's offset in the final stream;
  4. variable-size gather of the compressed bytes to rank 0 (batched isend / irecv: NCCL has no gatherv);
  5. rank 0 folds the per-rank Adler-32 values with adler32_combine and appends `03 00` + the big-endian checksum.

Inflate shards the same way (last sentence): the segment index of the stream is cut into contiguous ranges of
segments = contiguous compressed-offset ranges; every rank decodes its range, an all_gather of (decoded bytes, Adler-32,
Q14 carry, last byte) gives every rank its output offset and the stream's checksum (adler32_combine fold, checked against
the trailer), and the decoded bytes are optionally gathered to rank 0.

`engine` is anything with the Engine.deflate_segments_device / inflate_segments_device signature (tests drive the host logic with a CPU
stand-in built on the oracle; the product always passes zlibstream_b200.Engine).
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


_pinned = {}


def _to_host(t: torch.Tensor):
    """Small exchange tables go to the host through a reused pinned buffer: an asynchronous copy on the current stream and one
    stream synchronisation instead of the pageable, device-wide `.cpu()`."""
    if not t.is_cuda:
        return t.tolist()
    key = (t.shape, t.dtype)
    buf = _pinned.get(key)
    if buf is None:
        buf = _pinned[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return buf.tolist()


def shard_segments(nseg: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) range of segments of `rank`."""
    base, rem = divmod(nseg, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def exclusive_scan(sizes: List[int]) -> List[int]:
    out, acc = [], 0
    for s in sizes:
        out.append(acc)
        acc += s
    return out


def combine_adlers(adlers: List[int], lengths: List[int], combine) -> int:
    acc = 1
    for a, n in zip(adlers, lengths):
        acc = combine(acc, a, n)
    return acc


class ShardedDeflateResult:
    __slots__ = ("local_len", "local_adler", "sizes", "offsets", "total_len", "adler", "stream")

    def __init__(self):
        self.local_len = 0
        self.local_adler = 1
        self.sizes: List[int] = []
        self.offsets: List[int] = []
        self.total_len = 0
        self.adler = 1
        self.stream: Optional[torch.Tensor] = None   # rank 0 only, when gathered


def deflate_sharded(engine, combine, local_in: torch.Tensor, seg_size: int, level: int, strategy: int, hash_variant: int,
                    first_segment_index: int, local_out: torch.Tensor, seg_end: Optional[torch.Tensor] = None,
                    gather_to: Optional[torch.Tensor] = None, halo: Optional[torch.Tensor] = None,
                    group=None) -> ShardedDeflateResult:
    """Compresses this rank's shard and runs the exchange step.

    local_in / local_out / seg_end / gather_to / halo are tensors on this rank's device (uint8 except seg_end: int64
    viewed as uint64 offsets).  first_segment_index = stream index of this rank's first segment.  gather_to (rank 0
    only, may be None to skip the gather) receives header + all segments + trailer.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    res = ShardedDeflateResult()
    n = local_in.numel()
    out_len, adler = engine.deflate_segments_device(
        local_in.data_ptr(), n, seg_size, level, strategy, hash_variant, first_segment_index,
        local_out.data_ptr(), local_out.numel(), seg_end.data_ptr() if seg_end is not None else 0,
        header=(first_segment_index == 0), trailer=False, halo_ptr=halo.data_ptr() if halo is not None else 0)
    res.local_len, res.local_adler = out_len, adler
    if world == 1:
        res.sizes, res.offsets = [out_len], [0]
        res.adler = combine(1, adler, n)
        res.total_len = out_len + len(final_block(level)) + 4
        if gather_to is not None:
            gather_to[:out_len].copy_(local_out[:out_len])
            _write_trailer(gather_to, out_len, res.adler, level)
            res.stream = gather_to[:res.total_len]
        return res
    # 2. all_gather of per-rank (compressed bytes, input bytes, adler)
    # rank 0 decides whether the stream is gathered; the flag travels with the sizes so that every rank agrees
    mine = torch.tensor([out_len, n, adler, 1 if gather_to is not None else 0], dtype=torch.int64, device=local_in.device)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    table = _to_host(torch.stack(allv))
    gather_on = bool(table[0][3])
    res.sizes = [int(r[0]) for r in table]
    lengths = [int(r[1]) for r in table]
    adlers = [int(r[2]) for r in table]
    # 3. exclusive scan -> offsets
    res.offsets = exclusive_scan(res.sizes)
    res.adler = combine_adlers(adlers, lengths, combine)
    res.total_len = sum(res.sizes) + len(final_block(level)) + 4
    # 4. variable-size gather to rank 0
    if gather_on:
        _gather_to_rank0(rank, world, local_out, out_len, res.sizes, res.offsets, gather_to, group)
    # 5. trailer on rank 0
    if rank == 0 and gather_to is not None:
        _write_trailer(gather_to, sum(res.sizes), res.adler, level)
        res.stream = gather_to[:res.total_len]
    return res


def _gather_to_rank0(rank, world, local, local_len, sizes, offsets, gather_to, group):
    """Variable-size gather (NCCL has no gatherv): batched isend / irecv into rank 0's buffer at the scanned offsets."""
    ops = []
    if rank == 0:
        gather_to[:local_len].copy_(local[:local_len])
        for r in range(1, world):
            if sizes[r]:
                ops.append(dist.P2POp(dist.irecv, gather_to[offsets[r]:offsets[r] + sizes[r]], r, group))
    elif local_len:
        ops.append(dist.P2POp(dist.isend, local[:local_len], 0, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def shard_index(seg_ends: Sequence[int], header_len: int, world: int, rank: int) -> Tuple[int, int, int, int]:
    """This rank's share of a stream's segment index: (first segment, one past the last segment, first compressed byte,
    one past the last compressed byte).  seg_ends are the end offsets of the segments inside the stream (as written by
    zsb_deflate_segments or found by zsb_build_segment_index).  The header_len bytes in front of segment 0 belong to the
    slice of the rank that holds segment 0 (it passes has_header=True to inflate_sharded when header_len == 2)."""
    b, e = shard_segments(len(seg_ends), world, rank)
    lo = int(seg_ends[b - 1]) if b else 0
    hi = int(seg_ends[e - 1]) if e > b else max(lo, header_len if b == 0 else lo)
    return b, e, lo, hi


class ShardedInflateResult:
    __slots__ = ("local_len", "local_adler", "lengths", "offsets", "total_len", "adler", "carry", "out")

    def __init__(self):
        self.local_len = 0
        self.local_adler = 1
        self.lengths: List[int] = []
        self.offsets: List[int] = []
        self.total_len = 0
        self.adler = 1
        self.carry = 0
        self.out: Optional[torch.Tensor] = None   # rank 0 only, when gathered


def inflate_sharded(engine, combine, local_comp: torch.Tensor, local_ends: torch.Tensor, seg_size: int,
                    local_out: torch.Tensor, has_header: bool, expect_adler: Optional[int] = None,
                    gather_to: Optional[torch.Tensor] = None, group=None) -> ShardedInflateResult:
    """Decodes this rank's contiguous range of segments and runs the exchange step.

    local_comp = the stream's bytes [lo, hi) of shard_index (rank 0's slice starts with the 2-byte header when has_header);
    local_ends = int64 end offsets of this rank's segments relative to local_comp; local_out receives the decoded bytes.
    expect_adler = the stream's trailer value (every rank raises "incorrect data check" if the fold differs).
    gather_to (rank 0; None = leave the output sharded on the GPUs) receives the whole decoded stream.
    """
    from .options import ZlibStreamException
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    res = ShardedInflateResult()
    nseg = int(local_ends.numel())
    out_len, adler = engine.inflate_segments_device(local_comp.data_ptr(), int(local_comp.numel()), local_ends.data_ptr(), nseg,
                                                    seg_size, local_out.data_ptr(), int(local_out.numel()),
                                                    header=has_header, trailer=False, continuation=rank > 0)
    carry = int(engine.last_inflate_carry())
    res.local_len, res.local_adler, res.carry = out_len, adler, carry
    if world == 1:
        res.lengths, res.offsets, res.total_len, res.adler = [out_len], [0], out_len, adler
    else:
        # (decoded bytes, adler, Q14 carry of my first segment, my last byte, rank 0's gather flag)
        mine = torch.tensor([out_len, adler, carry, 0, 1 if gather_to is not None else 0], dtype=torch.int64, device=local_out.device)
        if out_len:
            mine[3] = local_out[out_len - 1]
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine, group=group)
        table = _to_host(torch.stack(allv))
        res.lengths = [int(r[0]) for r in table]
        adlers = [int(r[1]) for r in table]
        carries = [int(r[2]) for r in table]
        if any(carries[1:]):
            # Q14: a shard that opens with a run of the byte before it takes that byte from the nearest earlier shard whose
            # last byte is its own (a shard that is one carried run from end to end passes the byte on)
            if rank > 0 and carry:
                src = next((r for r in range(rank - 1, -1, -1) if res.lengths[r] > carries[r] or r == 0), 0)
                adler = engine.inflate_fill_carry(int(table[src][3]))
                res.local_adler = adler
            fixed = torch.tensor([adler], dtype=torch.int64, device=local_out.device)
            allf = [torch.empty_like(fixed) for _ in range(world)]
            dist.all_gather(allf, fixed, group=group)
            adlers = [int(t.item()) for t in allf]
        res.offsets = exclusive_scan(res.lengths)
        res.total_len = sum(res.lengths)
        res.adler = combine_adlers(adlers, res.lengths, combine)
        if bool(table[0][4]):
            _gather_to_rank0(rank, world, local_out, out_len, res.lengths, res.offsets, gather_to, group)
    if rank == 0 and gather_to is not None:
        if world == 1:
            gather_to[:out_len].copy_(local_out[:out_len])
        res.out = gather_to[:res.total_len]
    if expect_adler is not None and (res.adler & 0xFFFFFFFF) != (expect_adler & 0xFFFFFFFF):
        raise ZlibStreamException("inflating: incorrect data check")   # Inflate.cs:335-342
    return res


def final_block(level: int) -> bytes:
    """What Finish appends after the last full flush: the empty static block `03 00`; at level 0 the empty STORED
    block `01 00 00 FF FF` (Deflate.Stored.cs:80 with eof)."""
    return b"\x01\x00\x00\xff\xff" if level == 0 else b"\x03\x00"


def _write_trailer(buf: torch.Tensor, at: int, adler: int, level: int) -> int:
    raw = final_block(level) + bytes([(adler >> 24) & 0xFF, (adler >> 16) & 0xFF, (adler >> 8) & 0xFF, adler & 0xFF])
    buf[at:at + len(raw)].copy_(torch.tensor(list(raw), dtype=torch.uint8, device=buf.device))
    return len(raw)
