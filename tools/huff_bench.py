"""Huffman-stage timing (development tool): python tools/huff_bench.py [mib]  -> lz77 / huffman kernel ms for text L1 and
PNG-like HuffmanOnly / Rle (where the Huffman stage is the larger half)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools"); sys.path.insert(0, "tests")
import torch
import bench
from zlibstream_b200 import Engine

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mib << 20
eng = Engine(0)
dev = torch.device("cuda", 0)
for name, gen, level, strategy in (("text L1", "text", 1, 0), ("png HuffmanOnly", "png", 6, 2), ("png Rle", "png", 6, 3)):
    d = bench.device_input(gen, n, 0, dev)
    cap = eng.deflate_bound(n, 65536)
    out = torch.empty(cap, dtype=torch.uint8, device=dev)
    ends = torch.empty(n // 65536, dtype=torch.int64, device=dev)
    best = None
    for _ in range(3):
        eng.deflate_segments_device(d.data_ptr(), n, 65536, level, strategy, 0, 0, out.data_ptr(), cap, ends.data_ptr(), trailer=False)
        k = eng.last_kernel_ms()
        if best is None or k["huffman"] < best["huffman"]:
            best = k
    print(f"{name:18s} lz77 {best['lz77']:8.3f} ms  huffman {best['huffman']:7.3f} ms  -> {n / (best['lz77'] + best['huffman']) / 1e6:7.2f} GB/s")
eng.close()
