"""Inflate-only timing for profiling: python tools/inflate_prof.py [size_mib] [reps] [text|mixed|png] [segment_kib] [level] [strategy]
ZSB_INFLATE_KERNEL=warp|lanes|lanes_small forces a kernel."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools"); sys.path.insert(0, "tests")
import torch
import bench
from zlibstream_b200 import Engine

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else "text"
seg = (int(sys.argv[4]) if len(sys.argv) > 4 else 64) << 10
level = int(sys.argv[5]) if len(sys.argv) > 5 else 1
strategy = int(sys.argv[6]) if len(sys.argv) > 6 else 0
n = mib << 20
dev = torch.device("cuda", 0)
data = bench.device_input(kind, n, 0, dev)
eng = Engine(0)
cap = eng.deflate_bound(n, seg)
out = torch.empty(cap, dtype=torch.uint8, device="cuda")
ends = torch.empty(n // seg, dtype=torch.int64, device="cuda")
clen, _ = eng.deflate_segments_device(data.data_ptr(), n, seg, level, strategy, 0, 0, out.data_ptr(), cap, ends.data_ptr(), trailer=False)
dec = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    eng.inflate_segments_device(out.data_ptr(), clen, ends.data_ptr(), n // seg, seg, dec.data_ptr(), n, header=True, trailer=False)
    ms = eng.last_kernel_ms()["inflate"]
    print(f"inflate {kind} L{level} S{strategy} seg {seg >> 10} KiB, {mib} MiB ({n // seg} segments): {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s  ok={bool(torch.equal(dec, data))}")
eng.close()
