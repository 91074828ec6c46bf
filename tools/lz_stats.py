"""Prints the LZ77 kernel's diagnostics counters (ZSB_LZ_STATS=1) for one bench-shaped input.

Usage (under gpurun):  ZSB_LZ_STATS=1 python tools/lz_stats.py [size_mib] [level] [kind]
"""
import os
import struct
import sys

os.environ.setdefault("ZSB_LZ_STATS", "1")
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
sys.path.insert(0, "tests")

import torch  # noqa: E402

import synth  # noqa: E402
from zlibstream_b200 import Engine  # noqa: E402

NAMES = ["batches", "consumed", "dirty_stops", "nact_sum", "cyc_spec_links/insert", "cyc_rounds/search", "cyc_replay", "cyc_prod_wait",
         "matches/cyc_replay_wait", "x9", "cyc_fill", "cyc_parse", "x12", "x13", "serial_inserts", "rounds"]


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    level = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    kind = sys.argv[3] if len(sys.argv) > 3 else "text"
    n = mib << 20
    if kind == "text":
        data = synth.text_like(n, seed=1, device="cuda")
    else:
        import corpus_data
        if kind == "blob":
            c = corpus_data.load_corpus()
            b = b"".join(c[k] for k in sorted(c))
            blob = (b * (n // len(b) + 1))[:n]
        else:
            blob = getattr(corpus_data, kind)(n)
        data = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
    eng = Engine(0)
    cap = eng.deflate_bound(n, 65536)
    out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    ends = torch.empty(n // 65536 + 1, dtype=torch.int64, device="cuda")
    eng.deflate_segments_device(data.data_ptr(), n, 65536, level, 0, 0, 0, out.data_ptr(), cap, ends.data_ptr())
    before = struct.unpack("<16Q", eng.debug_fetch(6, 0, 128))
    res = eng.deflate_segments_device(data.data_ptr(), n, 65536, level, 0, 0, 0, out.data_ptr(), cap, ends.data_ptr())
    after = struct.unpack("<16Q", eng.debug_fetch(6, 0, 128))
    d = [a - b for a, b in zip(after, before)]
    print("kernel ms", eng.last_kernel_ms(), "result", res)
    for k, v in zip(NAMES, d):
        print(f"  {k:16s} {v:14d}  per-byte {v / n:10.4f}  per-batch {v / max(1, d[0]):10.2f}")
    eng.close()


if __name__ == "__main__":
    main()
