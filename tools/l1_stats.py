"""Diagnostics counters of the two level-1 kernels (ZSB_LZ_STATS=1): cycles of the linker warp, the walker warps and the
stage / slide phase per 32 KiB chunk of zsb_lz77_lists.cu.

Usage (under gpurun):  python tools/l1_stats.py [size_mib]
"""
import os
import struct
import sys

os.environ.setdefault("ZSB_LZ_STATS", "1")
sys.path.insert(0, ".")
sys.path.insert(0, "tools")

import synth  # noqa: E402
from zlibstream_b200 import Engine  # noqa: E402

NAMES = {16: "chunks", 17: "cyc_chunk_total", 18: "cyc_linker", 19: "cyc_linker_throttled", 20: "cyc_walkers_wait(sum)",
         21: "cyc_walkers(sum)", 22: "cyc_stage_slide", 23: "cyc_linker_waits_mail", 24: "groups", 25: "evals", 26: "steps", 27: "slow_searches", 28: "general_groups"}


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n = mib << 20
    data = synth.text_like(n, seed=1, device="cuda")
    eng = Engine(0)
    cap = eng.deflate_bound(n, 65536)
    import torch
    out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    ends = torch.empty(n // 65536 + 1, dtype=torch.int64, device="cuda")
    eng.deflate_segments_device(data.data_ptr(), n, 65536, 1, 0, 0, 0, out.data_ptr(), cap, ends.data_ptr())
    before = struct.unpack("<32Q", eng.debug_fetch(6, 0, 256))
    res = eng.deflate_segments_device(data.data_ptr(), n, 65536, 1, 0, 0, 0, out.data_ptr(), cap, ends.data_ptr())
    after = struct.unpack("<32Q", eng.debug_fetch(6, 0, 256))
    d = [a - b for a, b in zip(after, before)]
    print("kernel ms", eng.last_kernel_ms(), "result", res)
    chunks = max(1, d[16])
    for k, name in NAMES.items():
        print(f"  {name:24s} {d[k]:16d}  per chunk {d[k] / chunks:12.1f}  per byte {d[k] / n:8.3f}")
    eng.close()


if __name__ == "__main__":
    main()
