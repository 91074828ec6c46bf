"""Inflate-only timing for profiling: python tools/inflate_prof.py [size_mib] [reps]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import torch
import synth
from zlibstream_b200 import Engine

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = mib << 20
seg = 65536
data = synth.text_like(n, seed=1, device="cuda")
eng = Engine(0)
cap = eng.deflate_bound(n, seg)
out = torch.empty(cap, dtype=torch.uint8, device="cuda")
ends = torch.empty(n // seg, dtype=torch.int64, device="cuda")
clen, _ = eng.deflate_segments_device(data.data_ptr(), n, seg, 1, 0, 0, 0, out.data_ptr(), cap, ends.data_ptr(), trailer=False)
dec = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    eng.inflate_segments_device(out.data_ptr(), clen, ends.data_ptr(), n // seg, seg, dec.data_ptr(), n, header=True, trailer=False)
    ms = eng.last_kernel_ms()["inflate"]
    print(f"inflate {mib} MiB: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s  ok={bool(torch.equal(dec, data))}")
eng.close()
