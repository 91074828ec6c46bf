"""Multi-GPU parity check (run under torchrun on a GPU box; test infrastructure: it uses the oracle):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/dist_parity.py
Every rank compresses its contiguous range of segments, the exchange step gathers to rank 0, and rank 0 compares the gathered
stream byte for byte with the oracle's single-stream output for the same input (level 1, level 6, Rle with its one-byte
halo across ranks), then inflates it on the device without a side-car.  The gathered stream is then broadcast, every rank
decodes its contiguous range of the segment index (inflate_sharded: Q14 carries cross ranks, Adler-32 folds with
adler32_combine against the trailer) and rank 0 compares the gathered output with the input."""
import os
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch
import torch.distributed as dist

import corpus_data
from zlibstream_b200 import Engine, adler32_combine
from zlibstream_b200.dist import deflate_sharded, inflate_sharded, shard_index, shard_segments


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    seg = 65536
    data = corpus_data.text_like(9 * seg, seed=5) + bytes(3 * seg) + corpus_data.png_like(5 * seg + 4321, seed=6)
    nseg = (len(data) + seg - 1) // seg
    s0, s1 = shard_segments(nseg, world, rank)
    lo, hi = s0 * seg, min(len(data), s1 * seg)
    full = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    eng = Engine(local)
    ok_all = True
    for level, strategy in ((1, 0), (6, 0), (6, 3), (6, 2)):
        local_in = full[lo:hi].contiguous()
        halo = full[lo - 32768:lo].contiguous() if lo >= 32768 else None
        cap = eng.deflate_bound(hi - lo, seg)
        out = torch.empty(cap, dtype=torch.uint8, device=dev)
        ends = torch.empty(max(s1 - s0, 1), dtype=torch.int64, device=dev)
        gather = torch.empty(eng.deflate_bound(len(data), seg) + 64, dtype=torch.uint8, device=dev) if rank == 0 else \
            torch.empty(1, dtype=torch.uint8, device=dev)
        res = deflate_sharded(eng, adler32_combine, local_in, seg, level, strategy, 0, s0, out, ends, gather_to=gather, halo=halo)
        if rank == 0:
            import oracle_lib
            want, want_ends, want_adler = oracle_lib.load().deflate_segments(data, seg, level, strategy)
            got = bytes(res.stream.cpu().numpy())
            ok = got == want and res.adler == want_adler
            dec, ad = eng.inflate_segments(got, None, seg, len(data))
            ok = ok and dec == data and ad == want_adler
            print(f"world {world} level {level} strategy {strategy}: {'IDENTICAL' if ok else 'DIFFERENT'} ({len(got)} bytes)", flush=True)
            ok_all = ok_all and ok
        dist.barrier()
        # ---- sharded inflate of that stream: index and stream to every rank, each decodes its compressed-offset range ----
        meta = torch.zeros(2, dtype=torch.int64, device=dev)
        if rank == 0:
            meta[0], meta[1] = res.total_len, res.adler
        dist.broadcast(meta, 0)
        total_len, want_adler = int(meta[0]), int(meta[1])
        stream = gather[:total_len].clone() if rank == 0 else torch.empty(total_len, dtype=torch.uint8, device=dev)
        dist.broadcast(stream, 0)
        # segment ends of the whole stream: every rank's local ends + its scanned offset
        offs = torch.tensor(res.offsets, dtype=torch.int64)
        my = (ends[:s1 - s0] + int(offs[rank])).to(torch.int64)
        cnt = torch.tensor([s1 - s0], dtype=torch.int64, device=dev)
        cnts = [torch.empty_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        pad = torch.zeros(max(int(c) for c in cnts), dtype=torch.int64, device=dev)
        pad[:s1 - s0] = my
        pads = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(pads, pad)
        idx = torch.cat([p_[:int(c)] for p_, c in zip(pads, cnts)]).cpu().tolist()
        b, e, clo, chi = shard_index(idx, 2, world, rank)
        local_comp = stream[clo:chi].contiguous()
        local_ends = torch.tensor([x - clo for x in idx[b:e]], dtype=torch.int64, device=dev)
        dec = torch.zeros((e - b) * seg, dtype=torch.uint8, device=dev)
        gout = torch.empty(len(data) + 16, dtype=torch.uint8, device=dev) if rank == 0 else torch.empty(1, dtype=torch.uint8, device=dev)
        ires = inflate_sharded(eng, adler32_combine, local_comp, local_ends, seg, dec, has_header=(b == 0), expect_adler=want_adler,
                               gather_to=gout)
        if rank == 0:
            oki = bytes(ires.out.cpu().numpy()) == data and ires.adler == want_adler
            print(f"world {world} level {level} strategy {strategy}: sharded inflate {'IDENTICAL' if oki else 'DIFFERENT'} "
                  f"(carry of rank shards folded: adler {ires.adler:08x})", flush=True)
            ok_all = ok_all and oki
        dist.barrier()
    eng.close()
    dist.destroy_process_group()
    if rank == 0 and not ok_all:
        sys.exit(1)


if __name__ == "__main__":
    main()
