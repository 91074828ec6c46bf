"""Where the host-buffer deflate call spends its time (development tool): python tools/e2e_probe.py [mib]"""
import ctypes, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import torch
import synth
from zlibstream_b200 import Engine, _native

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mib << 20
seg = 65536
eng = Engine(0)
data = synth.text_like(n, seed=1, device="cuda")
cap = eng.deflate_bound(n, seg)
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.copy_(data)
h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
h_ends = torch.empty(n // seg, dtype=torch.int64).pin_memory()
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_ends = torch.empty(n // seg, dtype=torch.int64, device="cuda")
ol, ad = ctypes.c_size_t(0), ctypes.c_uint32(0)
for name, args in (("device", (data.data_ptr(), d_out.data_ptr(), d_ends.data_ptr(), _native.ZSB_IN_DEVICE | _native.ZSB_OUT_DEVICE | _native.ZSB_OFFSETS_DEVICE)),
                   ("host", (h_in.data_ptr(), h_out.data_ptr(), h_ends.data_ptr(), 0))):
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = eng._lib.zsb_deflate_segments(eng._h, args[0], n, seg, 1, 0, 0, 0, None, args[3] | _native.ZSB_NO_TRAILER, args[1], cap,
                                           ctypes.byref(ol), args[2], ctypes.byref(ad))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        k = eng.last_kernel_ms()
        if it:
            print(f"{name:6s} wall {dt:7.2f} ms  kernels lz77 {k['lz77']:.2f} (lists {k['lz77_lists']:.2f} resolve {k['lz77_resolve']:.2f}) huffman {k['huffman']:.2f}  sum {k['lz77'] + k['huffman']:.2f}")
