"""Throughput of the BASELINE.json configurations other than the headline one (kernel time, inputs resident in HBM).
python tools/config_sweep.py [size_mib]  -> one JSON line per configuration.  Data: tests/corpus_data generators (host),
replicated on the device to the requested size; every deflate output is inflated back and compared."""
import json
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
sys.path.insert(0, "tools")
import torch

import corpus_data
import synth
from zlibstream_b200 import Engine

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mib << 20
eng = Engine(0)


def tile(host_bytes, n):
    t = torch.frombuffer(bytearray(host_bytes), dtype=torch.uint8).cuda()
    reps = (n + t.numel() - 1) // t.numel()
    return t.repeat(reps)[:n].contiguous()


def run(name, data, level, strategy, seg):
    n = data.numel()
    nseg = (n + seg - 1) // seg
    cap = eng.deflate_bound(n, seg)
    out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    ends = torch.empty(nseg, dtype=torch.int64, device="cuda")
    best = None
    for _ in range(2):
        clen, _ = eng.deflate_segments_device(data.data_ptr(), n, seg, level, strategy, 0, 0, out.data_ptr(), cap, ends.data_ptr(),
                                              trailer=False)
        ms = eng.last_kernel_ms()
        tot = ms["lz77"] + ms["huffman"]
        best = tot if best is None else min(best, tot)
        lz, hf = ms["lz77"], ms["huffman"]
    dec = torch.empty(n, dtype=torch.uint8, device="cuda")
    ims = None
    for _ in range(2):
        eng.inflate_segments_device(out.data_ptr(), clen, ends.data_ptr(), nseg, seg, dec.data_ptr(), n, header=True, trailer=False)
        t = eng.last_kernel_ms()["inflate"]
        ims = t if ims is None else min(ims, t)
    ok = bool(torch.equal(dec, data))
    print(json.dumps({"config": name, "level": level, "strategy": strategy, "segment_bytes": seg, "mib": n >> 20,
                      "ratio": round(clen / n, 4), "deflate_gbs": round(n / best / 1e6, 2), "lz77_ms": round(lz, 2),
                      "huffman_ms": round(hf, 2), "inflate_gbs": round(n / ims / 1e6, 2), "round_trip_ok": ok}), flush=True)


text = synth.text_like(n, seed=1, device="cuda")
mixed = tile(corpus_data.mixed_entropy(32 << 20, seed=2), n)
png = tile(corpus_data.png_like(32 << 20, seed=3), n)
run("text-like", text, 1, 0, 65536)
run("text-like", text, 6, 0, 65536)
run("mixed-entropy", mixed, 1, 0, 262144)
run("mixed-entropy", mixed, 6, 0, 262144)
run("mixed-entropy", mixed, 9, 0, 262144)
run("png-like", png, 6, 1, 65536)
run("png-like", png, 6, 2, 65536)
run("png-like", png, 6, 3, 65536)
eng.close()
