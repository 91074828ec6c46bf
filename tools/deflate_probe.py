"""Device-resident deflate timing for any level / input kind (development tool; the command ncu profiles):
   python tools/deflate_probe.py [mib] [level] [text|mixed|png] [segment_kib] [strategy]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools"); sys.path.insert(0, "tests")
import torch
import bench
from zlibstream_b200 import Engine

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
kind = sys.argv[3] if len(sys.argv) > 3 else "text"
seg = (int(sys.argv[4]) if len(sys.argv) > 4 else 64) << 10
strategy = int(sys.argv[5]) if len(sys.argv) > 5 else 0
n = mib << 20
eng = Engine(0)
dev = torch.device("cuda", 0)
d = bench.device_input(kind, n, 0, dev)
cap = eng.deflate_bound(n, seg)
out = torch.empty(cap, dtype=torch.uint8, device=dev)
ends = torch.empty(n // seg, dtype=torch.int64, device=dev)
for _ in range(3):
    clen, _ad = eng.deflate_segments_device(d.data_ptr(), n, seg, level, strategy, 0, 0, out.data_ptr(), cap, ends.data_ptr(), trailer=False)
    k = eng.last_kernel_ms()
    print(f"{kind} L{level} S{strategy} seg {seg >> 10} KiB {mib} MiB: lz77 {k['lz77']:.2f} ms huffman {k['huffman']:.2f} ms -> "
          f"{n / (k['lz77'] + k['huffman']) / 1e6:.2f} GB/s, ratio {clen / n:.4f}")
eng.close()
