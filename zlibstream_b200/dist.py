"""Multi-GPU segmented deflate: one process per GPU, segments sharded contiguously across ranks.

The reference has no multi-process code at all; this is the exchange step the north star
describes for full-flush segments, which are independent units of work:

  1. every rank compresses its contiguous range of segments with the device engine (no collective on the data path);
     only stream segment 0 (rank 0's first) is "first" (header + 510-byte first chunk);
  2. all_gather of (compressed bytes, input bytes, Adler-32) per rank  -- NCCL over NVLink on GPUs, gloo in CPU tests;
  3. exclusive scan of the compressed sizes -> every rank's offset in the final stream;
  4. variable-size gather of the compressed bytes to rank 0 (batched isend / irecv: NCCL has no gatherv);
  5. rank 0 folds the per-rank Adler-32 values with adler32_combine and appends `03 00` + the big-endian checksum.

`engine` is anything with the Engine.deflate_segments_device signature (tests drive the host logic with a CPU
stand-in built on the oracle; the product always passes zlibstream_b200.Engine).
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


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
        res.total_len = out_len + 6
        if gather_to is not None:
            gather_to[:out_len].copy_(local_out[:out_len])
            _write_trailer(gather_to, out_len, res.adler)
            res.stream = gather_to[:res.total_len]
        return res
    # 2. all_gather of per-rank (compressed bytes, input bytes, adler)
    mine = torch.tensor([out_len, n, adler], dtype=torch.int64, device=local_in.device)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    table = torch.stack(allv).cpu().tolist()
    res.sizes = [int(r[0]) for r in table]
    lengths = [int(r[1]) for r in table]
    adlers = [int(r[2]) for r in table]
    # 3. exclusive scan -> offsets
    res.offsets = exclusive_scan(res.sizes)
    res.adler = combine_adlers(adlers, lengths, combine)
    res.total_len = sum(res.sizes) + 6
    # 4. variable-size gather to rank 0
    if gather_to is not None or rank != 0:
        ops = []
        if rank == 0:
            if gather_to is not None:
                gather_to[:out_len].copy_(local_out[:out_len])
                for r in range(1, world):
                    if res.sizes[r]:
                        ops.append(dist.P2POp(dist.irecv, gather_to[res.offsets[r]:res.offsets[r] + res.sizes[r]], r, group))
        elif out_len and _wants_gather(gather_to, group):
            ops.append(dist.P2POp(dist.isend, local_out[:out_len], 0, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
    # 5. trailer on rank 0
    if rank == 0 and gather_to is not None:
        _write_trailer(gather_to, sum(res.sizes), res.adler)
        res.stream = gather_to[:res.total_len]
    return res


def _wants_gather(gather_to, group) -> bool:
    # non-zero ranks cannot see rank 0's gather_to; by convention they pass a 1-element tensor (or anything not None)
    # when the gather is on, None when it is off
    return gather_to is not None


def _write_trailer(buf: torch.Tensor, at: int, adler: int):
    t = torch.tensor([0x03, 0x00, (adler >> 24) & 0xFF, (adler >> 16) & 0xFF, (adler >> 8) & 0xFF, adler & 0xFF],
                     dtype=torch.uint8, device=buf.device)
    buf[at:at + 6].copy_(t)
