"""Multi-GPU parity check (run under torchrun on a GPU box; test infrastructure: it uses the oracle):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/dist_parity.py
Every rank compresses its contiguous range of segments, the exchange step gathers to rank 0, and rank 0 compares the gathered
stream byte for byte with the oracle's single-stream output for the same input (level 1, level 6, Rle with its one-byte
halo across ranks), then inflates it on the device without a side-car."""
import os
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch
import torch.distributed as dist

import corpus_data
from zlibstream_b200 import Engine, adler32_combine
from zlibstream_b200.dist import deflate_sharded, shard_segments


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
    eng.close()
    dist.destroy_process_group()
    if rank == 0 and not ok_all:
        sys.exit(1)


if __name__ == "__main__":
    main()
