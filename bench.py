"""bench.py -- segmented DEFLATE / INFLATE throughput on B200 through libzsb200's C ABI (BASELINE.json configs 2-5).

  python bench.py --gpus 1 --steps 5 --warmup 3                  # config 2 (the headline): this repo's CUDA path
  python bench.py --config 3|4|5 ...                             # the other BASELINE.json configurations
  python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 # the reference's algorithm on the host cores (oracle)
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # one rank per GPU, segments sharded (weak scaling)

A step = one pass of the hot path over one batch of synthetic input:
  config 2  1 GiB of text-like data per GPU, level 1 (deflate_fast), 64 KiB full-flush segments            [deflate]
  config 3  2 GiB of mixed-entropy data per GPU (16 GiB over 8), level 6, 256 KiB segments (+ level 9 on a sub-sample)
  config 4  inflate of a full-flush-segmented stream, 2 GiB decoded per GPU at N > 1, 16 GiB on one GPU     [inflate]
  config 5  8 GiB of PNG-style filtered RGBA scanlines per GPU (64 GiB over 8): Filtered L6 (+ HuffmanOnly, Rle)
`value` = GB/s of uncompressed bytes with the input resident in HBM; `e2e` = the same call with pinned HOST buffers (copies
inside the timed region).  Rank 0 prints ONE JSON line.  The CPU legs time the oracle (a C restatement of the reference's
managed algorithm; the reference itself needs .NET) with native pinned threads (oracle/zso_bench.c).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries the ONE JSON line: keep NCCL's version banner (printed to stdout at NCCL_DEBUG=VERSION / WARN) out of it
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    del os.environ["NCCL_DEBUG"]

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

HASH = 0
UNIT = "GB/s"
STRATEGY_NAMES = {0: "Default", 1: "Filtered", 2: "HuffmanOnly", 3: "Rle", 4: "Fixed"}

CONFIGS = {
    2: dict(kind="deflate", gen="text", mib=1024, seg=65536, level=1, strategy=0, metric="segmented_deflate_throughput",
            workload="text-like synthetic, 1 GiB per GPU, level 1 (deflate_fast), 64 KiB full-flush segments"),
    3: dict(kind="deflate", gen="mixed", mib=2048, seg=262144, level=6, strategy=0, metric="segmented_deflate_throughput",
            workload="mixed-entropy synthetic, 2 GiB per GPU (16 GiB over 8 GPUs), level 6 (lazy deflate_slow), 256 KiB full-flush segments"),
    4: dict(kind="inflate", gen="text", mib=2048, seg=65536, level=1, strategy=0, metric="segmented_inflate_throughput",
            workload="inflate of a full-flush-segmented zlib stream (level-1 text stream, 64 KiB segments), 2 GiB decoded per GPU "
                     "(16 GiB over 8 GPUs; 16 GiB on the one GPU at N = 1)"),
    5: dict(kind="deflate", gen="png", mib=8192, seg=65536, level=6, strategy=1, metric="segmented_deflate_throughput",
            workload="PNG-style filtered RGBA scanlines, 8 GiB per GPU (64 GiB over 8 GPUs), Z_FILTERED level 6, 64 KiB full-flush segments"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--size-mib", type=int, default=0, help="input MiB per GPU and step (0 = the configuration's size)")
    ap.add_argument("--cpu-sample-mib", type=int, default=0, help="0 = sized for ~10 s of CPU work")
    ap.add_argument("--no-inflate", action="store_true")
    ap.add_argument("--no-inflate-large", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--gather", action="store_true", help="config 4: also gather the decoded bytes to rank 0 inside the timed region")
    a = ap.parse_args()
    a.cfg = dict(CONFIGS[a.config])
    if a.size_mib:
        a.cfg["mib"] = a.size_mib
    a.size_mib = a.cfg["mib"]
    return a


def workload_config(args, world):
    c = args.cfg
    n = c["mib"] << 20
    return {"workload": c["workload"], "baseline_config": args.config, "bytes_per_gpu": n, "segment_bytes": c["seg"], "level": c["level"],
            "strategy": STRATEGY_NAMES[c["strategy"]], "hash": "crc32c", "segments_per_gpu": n // c["seg"],
            "sharding": f"segments/{world} ranks",
            "l2_policy": f"input ({c['mib']} MiB) and scratch are larger than the 126 MB L2"}


# ---------------------------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------------------------

def host_bytes(gen, nbytes, seed):
    import corpus_data
    if gen == "text":
        import synth
        return synth.text_like(nbytes, seed=seed, device="cpu", chunk=16 << 20).numpy().tobytes()
    base = (corpus_data.mixed_entropy if gen == "mixed" else corpus_data.png_like)(min(nbytes, 32 << 20), seed=seed)
    return (base * ((nbytes + len(base) - 1) // len(base)))[:nbytes]


def device_input(gen, nbytes, rank, dev):
    """uint8 tensor of nbytes on dev: text from the device generator; mixed / PNG-style = 32 MiB of the host generator tiled
    (matches reach back at most 32 KiB, so the period does not show)."""
    import torch
    if gen == "text":
        import synth
        return synth.text_like(nbytes, seed=1 + rank, device=dev)
    seed = (2 if gen == "mixed" else 3) + 10 * rank
    t = torch.frombuffer(bytearray(host_bytes(gen, min(nbytes, 32 << 20), seed)), dtype=torch.uint8).to(dev)
    return t.repeat((nbytes + t.numel() - 1) // t.numel())[:nbytes].contiguous()


# ---------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle (a C port of the reference's managed algorithm) and system zlib on the host cores, native threads
# ---------------------------------------------------------------------------------------------------------------

def cpu_deflate(sample: bytes, cfg, threads: int, target_s: float):
    """-> dict(value GB/s, seconds, passes, ratio, per_core_mbs): passes over `sample` until ~target_s of wall time."""
    import oracle_lib
    orc = oracle_lib.load()
    sec, comp = orc.bench_deflate_mt(sample, cfg["seg"], cfg["level"], cfg["strategy"], HASH, threads, 1)   # calibration, counted
    total, passes = sec, 1
    reps = int(min(512, max(0, round(target_s / max(sec, 1e-3)) - 1)))
    if reps:
        sec2, comp = orc.bench_deflate_mt(sample, cfg["seg"], cfg["level"], cfg["strategy"], HASH, threads, reps)
        total += sec2
        passes += reps
    gbs = len(sample) * passes / total / 1e9
    return {"value": gbs, "seconds": total, "passes": passes, "ratio": comp / len(sample), "per_core_mbs": gbs * 1e3 / threads}


def cpu_inflate(stream: bytes, ends, first_offset, seg, decoded_bytes, which, threads, target_s):
    import oracle_lib
    orc = oracle_lib.load()
    sec, dec = orc.bench_inflate_mt(stream, ends, first_offset, seg, which, threads, 1)
    assert dec == decoded_bytes, (dec, decoded_bytes)
    total, passes = sec, 1
    reps = int(min(512, max(0, round(target_s / max(sec, 1e-3)) - 1)))
    if reps:
        sec2, _ = orc.bench_inflate_mt(stream, ends, first_offset, seg, which, threads, reps)
        total += sec2
        passes += reps
    gbs = decoded_bytes * passes / total / 1e9
    return {"value": gbs, "seconds": total, "passes": passes, "per_core_mbs": gbs * 1e3 / threads}


def cpu_single_stream(sample: bytes, cfg):
    """What the reference itself does: ONE stream, one thread.  (a) the same full-flush segments, sequentially; (b) one Write,
    NoFlush (the reference's benchmark call pattern) -> the unsegmented ratio the north star asks to print beside ours."""
    import oracle_lib
    orc = oracle_lib.load()
    sec, comp = orc.bench_deflate_mt(sample, cfg["seg"], cfg["level"], cfg["strategy"], HASH, 1, 1)
    t0 = time.perf_counter()
    one, _ = orc.deflate_oneshot(sample, cfg["level"], cfg["strategy"], HASH)
    t1 = time.perf_counter() - t0
    return {"sample_mib": len(sample) >> 20, "segmented_1core_mbs": len(sample) / sec / 1e6, "ratio_segmented": comp / len(sample),
            "unsegmented_1core_mbs": len(sample) / t1 / 1e6, "ratio_unsegmented": len(one) / len(sample),
            "note": "C oracle, one thread; the managed reference publishes 73 / 20 MB/s (corpus, level 1 / 6, i7-8650U; BASELINE.md 2.2)"}


def run_reference(args):
    """--impl reference: the reference's own (CPU) algorithm for the same metric and config.  The managed reference cannot
    run here (no .NET); its C restatement in oracle/ is timed with every host thread (inflate: system zlib)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.cfg
    cores = os.cpu_count() or 1
    per_core = {1: 40, 6: 14, 9: 2}.get(cfg["level"], 14) if cfg["strategy"] in (0, 1) else 150   # rough MB/s, sizes the sample
    sample_mib = args.cpu_sample_mib or int(max(16, min(cfg["mib"], per_core * cores * 2 / 4)))
    sample_mib = max(cfg["seg"] * cores >> 20, sample_mib)
    sample = host_bytes(cfg["gen"], sample_mib << 20, 1)
    import oracle_lib
    orc = oracle_lib.load()
    target = max(2.0, 20.0 / max(1, args.steps))
    if cfg["kind"] == "deflate":
        for _ in range(args.warmup):
            orc.bench_deflate_mt(sample[: (8 << 20)], cfg["seg"], cfg["level"], cfg["strategy"], HASH, cores, 1)
        runs = [cpu_deflate(sample, cfg, cores, target) for _ in range(args.steps)]
        what = f"ratio {runs[-1]['ratio']:.3f}"
    else:
        stream, ends, _ = orc.deflate_segments_parallel(sample, cfg["seg"], cfg["level"], cfg["strategy"], HASH)
        for _ in range(args.warmup):
            orc.bench_inflate_mt(stream, ends[:64], 2, cfg["seg"], 1, cores, 1)
        runs = [cpu_inflate(stream, ends, 2, cfg["seg"], len(sample), 1, cores, target) for _ in range(args.steps)]
        what = "system zlib " + __import__("zlib").ZLIB_RUNTIME_VERSION + ", one segment at a time per thread"
    t_total = sum(r["seconds"] for r in runs)
    passes = sum(r["passes"] for r in runs)
    value = len(sample) * passes / t_total / 1e9
    line = {"impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "per_core_mbs": value * 1e3 / cores,
                             "sample": f"{passes} passes over {sample_mib} MiB of the same workload ({t_total:.1f} s of wall time in "
                                       f"total), {cores} native pinned threads, one contiguous segment range (own stream) per thread, {what}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nme in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------------------
# parity of the bench's own bytes with the oracle (outside every timed region)
# ---------------------------------------------------------------------------------------------------------------

def parity_vs_oracle(data, out, seg_end, cfg, first_seg, hdr_len, budget_mib):
    """sha256 of the device's compressed bytes for the first segments and for a slice from the middle of the bench input
    against the oracle's bytes for the same segments (a segment k >= 1 compresses to the same bytes whatever came
    before it, given its own 32 KiB of history: the oracle gets the preceding segment in front and its bytes are dropped)."""
    import oracle_lib
    orc = oracle_lib.load()
    seg = cfg["seg"]
    nseg = data.numel() // seg
    k = max(1, min(nseg // 2, (budget_mib << 20) // 2 // seg))
    ends = seg_end.cpu().tolist()
    checked = []
    for s0 in (0, nseg // 2):
        if s0 == 0 and first_seg == 0:     # the stream's first segments: header + segments, nothing dropped
            host, drop = data[:k * seg].cpu().numpy().tobytes(), 0
        elif s0 == 0:                      # a later rank's first segments: a filler segment in front, its bytes dropped
            host, drop = bytes(seg) + data[:k * seg].cpu().numpy().tobytes(), 1
        else:                              # a slice from the middle: the real preceding segment in front, its bytes dropped
            host, drop = data[(s0 - 1) * seg:(s0 + k) * seg].cpu().numpy().tobytes(), 1
        want, wends, _ = orc.deflate_segments_parallel(host, seg, cfg["level"], cfg["strategy"], HASH)
        wbytes = want[(wends[0] if drop else 2):wends[drop + k - 1]]
        g0 = ends[s0 - 1] if s0 else hdr_len
        gbytes = out[g0:ends[s0 + k - 1]].cpu().numpy().tobytes()
        ok = hashlib.sha256(gbytes).digest() == hashlib.sha256(wbytes).digest()
        checked.append({"first_segment": s0, "segments": k, "identical": ok})
    return {"identical": all(c["identical"] for c in checked), "slices": checked,
            "bytes_checked": 2 * k * seg, "how": "sha256 of the device stream vs the oracle's stream for the same segments"}


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------

def _hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "6.65 TB/s (of fallback, B200_PROFILING.md)"


def _traffic(kernel_key, n):
    """dram bytes per launch from this round's committed `ncu --set full` capture of the same command
    (profiles/r02_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per input byte, captured at bench size)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))[kernel_key]
        return t["dram_bytes_per_unit_byte"] * n
    except Exception:
        return None


class Ctx:
    pass


def run_ours(args):
    import torch
    import torch.distributed as dist
    from zlibstream_b200 import Engine

    c = Ctx()
    c.args, c.cfg = args, args.cfg
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: libzsb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    c.dev = torch.device("cuda", local_rank)
    if c.world > 1:
        dist.init_process_group("nccl", device_id=c.dev)
    c.eng = Engine(local_rank)
    c.stream = torch.cuda.Stream(device=c.dev)
    c.eng.set_stream(c.stream.cuda_stream)
    c.local_rank = local_rank

    def barrier():
        if c.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    c.barrier = barrier
    line = bench_deflate(c) if c.cfg["kind"] == "deflate" else bench_inflate(c)
    if c.rank == 0:
        print(json.dumps(line))
    if c.world > 1:
        dist.destroy_process_group()


def max_over_ranks(c, ms):
    import torch
    import torch.distributed as dist
    if c.world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=c.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


def bench_deflate(c):
    import torch
    from zlibstream_b200 import adler32_combine
    from zlibstream_b200.dist import deflate_sharded
    args, cfg, eng, dev, stream = c.args, c.cfg, c.eng, c.dev, c.stream
    seg, level, strategy = cfg["seg"], cfg["level"], cfg["strategy"]
    n = cfg["mib"] << 20
    nseg = n // seg
    data = device_input(cfg["gen"], n, c.rank, dev)
    cap = eng.deflate_bound(n, seg)
    out = torch.empty(cap, dtype=torch.uint8, device=dev)
    seg_end = torch.empty(nseg, dtype=torch.int64, device=dev)
    gather = torch.empty(cap * c.world + 64, dtype=torch.uint8, device=dev) if (c.world > 1 and c.rank == 0) else \
        (torch.empty(1, dtype=torch.uint8, device=dev) if c.world > 1 else None)
    first_seg = c.rank * nseg
    halo = None   # Rle looks one byte back across ranks; ranks here hold independent inputs (weak scaling), so no halo
    torch.cuda.synchronize()

    def step():
        with torch.cuda.stream(stream):
            return deflate_sharded(eng, adler32_combine, data, seg, level, strategy, HASH, first_seg, out, seg_end,
                                   gather_to=gather, halo=halo)

    for _ in range(max(args.warmup, 0)):
        res = step()
    c.barrier()
    launches0 = eng.launch_count
    sampler = ClockSampler(c.local_rank)
    if c.rank == 0:
        sampler.start()
    kms = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
    for _ in range(args.steps):
        res = step()
        kms.append(eng.last_kernel_ms())
    with torch.cuda.stream(stream):
        ev1.record(stream)
    c.barrier()
    elapsed_ms = max_over_ranks(c, ev0.elapsed_time(ev1))
    clocks = sampler.stop() if c.rank == 0 else None
    launches = eng.launch_count - launches0
    value = n * c.world * args.steps / (elapsed_ms / 1e3) / 1e9
    ratio = res.local_len / n
    hdr_len = 2 if c.rank == 0 else 0

    parity = None
    if not args.no_parity:
        budget = {1: 64, 6: 64, 9: 8}.get(level, 32)
        parity = parity_vs_oracle(data, out, seg_end, cfg, first_seg, hdr_len, budget)

    inflate = None
    if not args.no_inflate:
        inflate = inflate_of_stream(c, data, out, res.local_len, seg_end, nseg, ratio)
        if args.config == 2 and not args.no_inflate_large and cfg["mib"] >= 1024:
            inflate_large(c, inflate, data, out, res.local_len, seg_end, nseg, ratio)

    other = None
    if c.world == 1 and args.config == 2 and not args.no_other_configs:
        other = other_configs(c, data)
    extra = None
    if args.config == 3 and not args.no_other_configs:   # level 9 on a stated sub-sample (chain 4096)
        extra = [run_config(c, "mixed-entropy, level 9, 256 KiB segments (256 MiB sub-sample of this rank's input)",
                            data[:256 << 20].contiguous(), 9, 0, seg, parity_mib=8)]
    if args.config == 5 and not args.no_other_configs:
        sub = data[:1024 << 20].contiguous()
        extra = [run_config(c, "PNG-style, HuffmanOnly (1 GiB sub-sample of this rank's input)", sub, 6, 2, seg, parity_mib=32),
                 run_config(c, "PNG-style, Rle (1 GiB sub-sample of this rank's input)", sub, 6, 3, seg, parity_mib=32)]

    e2e = None
    if not args.no_e2e:
        e2e = e2e_deflate(c, data, n, cap, out, res.local_len)

    if c.rank != 0:
        return None
    peak, peak_src = _hbm_peak()
    lz = sum(k["lz77"] for k in kms) / len(kms)
    hf = sum(k["huffman"] for k in kms) / len(kms)
    la = sum(k["lz77_lists"] for k in kms) / len(kms)
    rb = sum(k["lz77_resolve"] for k in kms) / len(kms)
    algo_bytes = n * (1 + ratio)
    two = la > 0 and rb > 0
    kname = "lz77 stage = lz77_lists_kernel + lz77_resolve_kernel (one launch each per batch)" if two else "lz77_segments_kernel"
    achieved = algo_bytes / (lz / 1e3) / 1e9
    kernel_ms = {"lz77_stage": lz, "huffman_stage": hf}
    if two:
        kernel_ms.update({"lz77_lists_kernel": la, "lz77_resolve_kernel": rb})
    line = {"metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": c.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(args, c.world),
            "compression_ratio": ratio, "gpu_launches": launches, "kernel_ms": kernel_ms,
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": _traffic(f"deflate_config{args.config}", n),
                         "note": f"algorithmic bytes = input * (1 + ratio) = {algo_bytes:.0f} per step; peak = {peak_src}; "
                                 "the stage is issue / dependency-bound, not bandwidth-bound (DESIGN.md); traffic = dram bytes of "
                                 "this round's ncu capture at bench size (the level-1 path moves 32 B of candidate records per "
                                 "input byte between its two kernels)"},
            "clocks": clocks}
    if parity:
        line["parity_vs_oracle"] = parity["identical"]
        line["parity_detail"] = parity
    if inflate:
        line["inflate"] = inflate
    if e2e:
        line["e2e"] = e2e
    if other:
        line["other_configs"] = other
    if extra:
        line["same_config_other_settings"] = extra
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        per_core = {1: 40, 6: 14, 9: 2}.get(level, 14) if strategy in (0, 1) else 150
        sample_mib = args.cpu_sample_mib or int(max(16, min(cfg["mib"], per_core * cores / 2)))
        sample = data[: sample_mib << 20].cpu().numpy().tobytes()
        r = cpu_deflate(sample, cfg, cores, 12.0)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port", "per_core_mbs": r["per_core_mbs"],
                                "sample": f"{r['passes']} passes over the first {sample_mib} MiB of rank 0's input ({r['seconds']:.1f} s), "
                                          f"{cores} native pinned threads, one contiguous segment range (own stream) per thread; "
                                          "the C oracle restates the reference's managed algorithm (no .NET here)"}
        line["cpu_single_stream"] = cpu_single_stream(sample[: min(len(sample), 32 << 20)], cfg)
        line["ratio_unsegmented"] = line["cpu_single_stream"]["ratio_unsegmented"]
    return line


def inflate_of_stream(c, data, out, comp_len, seg_end, nseg, ratio):
    """Inflate of the stream just produced (bit-identical to the reference's, see tests): device-resident, without the
    side-car, end to end with host buffers, and the CPU figures (system zlib: one core sequential = what ZlibInputStream
    does, and all cores by segment)."""
    import ctypes
    import torch
    from zlibstream_b200 import _native
    args, cfg, eng, dev, stream = c.args, c.cfg, c.eng, c.dev, c.stream
    seg = cfg["seg"]
    n = data.numel()
    hdr = c.rank == 0
    dec = torch.empty(n, dtype=torch.uint8, device=dev)

    def timed(fn, reps):
        with torch.cuda.stream(stream):
            for _ in range(2):
                fn()
            c.barrier()
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            i0.record(stream)
            ks = []
            for _ in range(reps):
                fn()
                ks.append(eng.last_kernel_ms()["inflate"])
            i1.record(stream)
        c.barrier()
        return max_over_ranks(c, i0.elapsed_time(i1)) / reps, sum(ks) / len(ks)

    isteps = max(1, min(args.steps, 3))
    ms, kms = timed(lambda: eng.inflate_segments_device(out.data_ptr(), comp_len, seg_end.data_ptr(), nseg, seg, dec.data_ptr(), n,
                                                        header=hdr, trailer=False), isteps)
    ok = bool(torch.equal(dec, data))
    res = {"value": n * c.world / (ms / 1e3) / 1e9, "unit": UNIT, "ms_per_step": ms, "kernel_ms": kms, "round_trip_ok": ok,
           "steps": isteps, "roofline_frac": (n * (1 + ratio) / (kms / 1e3) / 1e9) / _hbm_peak()[0]}
    if c.world == 1:
        # without the side-car: marker scan + count-only decode + decode (zsb_build_segment_index inside the call)
        dec.zero_()
        ms2, _ = timed(lambda: eng.inflate_segments_device(out.data_ptr(), comp_len, 0, 0, seg, dec.data_ptr(), n, header=hdr,
                                                           trailer=False), 2)
        res["without_side_car"] = {"value": n / (ms2 / 1e3) / 1e9, "unit": UNIT, "ms_per_step": ms2, "round_trip_ok": bool(torch.equal(dec, data))}
        if not args.no_e2e:
            h_in = torch.empty(comp_len, dtype=torch.uint8).pin_memory()
            h_in.copy_(out[:comp_len])
            h_ends = seg_end.cpu().pin_memory()
            h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
            ol, ad = ctypes.c_size_t(0), ctypes.c_uint32(0)
            flags = (0 if hdr else _native.ZSB_NO_HEADER) | _native.ZSB_NO_TRAILER

            def call():
                rc = eng._lib.zsb_inflate_segments(eng._h, h_in.data_ptr(), comp_len, h_ends.data_ptr(), nseg, seg, flags,
                                                   h_out.data_ptr(), n, ctypes.byref(ol), ctypes.byref(ad))
                if rc not in (0, 1):
                    raise RuntimeError("e2e inflate failed: " + eng.message)

            ms3, _ = timed(call, 2)
            res["e2e"] = {"value": n / (ms3 / 1e3) / 1e9, "unit": UNIT, "ms_per_step": ms3, "h2d_bytes_per_step": comp_len + 8 * nseg,
                          "d2h_bytes_per_step": n, "api": "zsb_inflate_segments with pinned host buffers",
                          "round_trip_ok": bool(torch.equal(h_out.to(dev), data))}
    if c.rank == 0 and not args.no_cpu:
        cores = os.cpu_count() or 1
        smib = min(n >> 20, 256)
        sn = smib << 20
        k = sn // seg
        ends = seg_end[:k].cpu().tolist()
        hbytes = out[:ends[-1]].cpu().numpy().tobytes()
        one = cpu_inflate(hbytes, ends, 2, seg, sn, 1, 1, 4.0)
        allc = cpu_inflate(hbytes, ends, 2, seg, sn, 1, cores, 6.0)
        res["cpu_baseline"] = {"value": allc["value"], "unit": UNIT, "cores": cores, "kind": "system zlib " + __import__("zlib").ZLIB_RUNTIME_VERSION,
                               "per_core_mbs": allc["per_core_mbs"], "one_core_sequential_mbs": one["value"] * 1e3,
                               "sample": f"first {smib} MiB of rank 0's stream, {allc['passes']} passes, native pinned threads, one segment at "
                                         "a time per thread (BASELINE.md section 4: the reference's own inflater mis-builds dynamic tables)"}
    return res


def inflate_large(c, inflate, data, out, comp_len, seg_end, nseg, ratio):
    """Inflate at the sizes of BASELINE config 4: the stream body above 2 (N > 1: 2 GiB per GPU) or 2 and 16 times in a row on
    one GPU -- full-flush segments are independent, so the concatenation is a valid segmented body."""
    import torch
    eng, dev, stream, cfg = c.eng, c.dev, c.stream, c.cfg
    seg = cfg["seg"]
    n = data.numel()
    for reps in ((2, 16) if c.world == 1 else (2,)):
        hdr_len = 2 if c.rank == 0 else 0
        body = out[hdr_len:comp_len]
        big = body.repeat(reps)
        ends1 = seg_end - hdr_len
        big_ends = torch.cat([ends1 + r * body.numel() for r in range(reps)]).contiguous()
        big_dec = torch.empty(n * reps, dtype=torch.uint8, device=dev)
        with torch.cuda.stream(stream):
            bms = []
            for it in range(3):
                eng.inflate_segments_device(big.data_ptr(), big.numel(), big_ends.data_ptr(), nseg * reps, seg,
                                            big_dec.data_ptr(), n * reps, header=False, trailer=False)
                if it:
                    bms.append(eng.last_kernel_ms()["inflate"])
        torch.cuda.synchronize()
        okb = bool(torch.equal(big_dec[:n], data)) and bool(torch.equal(big_dec[-n:], data))
        bm = sum(bms) / len(bms)
        inflate["at_2gib_per_gpu" if reps == 2 else "at_16gib_one_gpu"] = {
            "value": n * reps / (bm / 1e3) / 1e9, "unit": UNIT, "kernel_ms": bm, "segments": nseg * reps, "round_trip_ok": okb,
            "scope": "this rank's kernel time (ranks are independent)",
            "roofline_frac": (n * reps * (1 + ratio) / (bm / 1e3) / 1e9) / _hbm_peak()[0]}
        del big, big_dec, big_ends


def e2e_deflate(c, data, n, cap, dev_out, dev_len):
    import ctypes
    import torch
    from zlibstream_b200 import _native
    args, cfg, eng, stream = c.args, c.cfg, c.eng, c.stream
    seg = cfg["seg"]
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.copy_(data)
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    nseg = n // seg
    h_ends = torch.empty(nseg, dtype=torch.int64).pin_memory()
    lib = eng._lib
    out_len = ctypes.c_size_t(0)
    adler = ctypes.c_uint32(0)
    flags = (0 if c.rank == 0 else _native.ZSB_NO_HEADER) | _native.ZSB_NO_TRAILER

    def call():
        rc = lib.zsb_deflate_segments(eng._h, h_in.data_ptr(), n, seg, cfg["level"], cfg["strategy"], HASH, c.rank * nseg, None, flags,
                                      h_out.data_ptr(), cap, ctypes.byref(out_len), h_ends.data_ptr(), ctypes.byref(adler))
        if rc != 0:
            raise RuntimeError("e2e deflate failed: " + eng.message)

    call()
    c.barrier()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            call()
        e1.record(stream)
    c.barrier()
    ms = max_over_ranks(c, max(e0.elapsed_time(e1), 0.0))
    wall = (time.perf_counter() - t0) * 1e3
    same = int(out_len.value) == dev_len and bool(torch.equal(h_out[:dev_len], dev_out[:dev_len].cpu()))
    return {"value": n * c.world * steps / (ms / 1e3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": n,
            "d2h_bytes_per_step": int(out_len.value) + 8 * nseg + 16, "ms_per_step": ms / steps, "wall_ms_per_step": wall / steps,
            "api": "zsb_deflate_segments with pinned host buffers", "bytes_equal_device_path": same}


def run_config(c, name, d, level, strategy, seg, parity_mib=0):
    """Deflate + inflate kernel throughput of one more setting (kernel time; context for the headline, not a bench line).
    Every output is inflated back and compared; the first segments and a middle slice are compared with the oracle."""
    import torch
    eng, dev = c.eng, c.dev
    nseg = (d.numel() + seg - 1) // seg
    cap = eng.deflate_bound(d.numel(), seg)
    out = torch.empty(cap, dtype=torch.uint8, device=dev)
    ends = torch.empty(nseg, dtype=torch.int64, device=dev)
    best = None
    for _ in range(2):
        clen, _ = eng.deflate_segments_device(d.data_ptr(), d.numel(), seg, level, strategy, HASH, 0, out.data_ptr(), cap,
                                              ends.data_ptr(), trailer=False)
        k = eng.last_kernel_ms()
        t = k["lz77"] + k["huffman"]
        best = t if best is None else min(best, t)
    dec = torch.empty(d.numel(), dtype=torch.uint8, device=dev)
    ims = None
    for _ in range(2):
        eng.inflate_segments_device(out.data_ptr(), clen, ends.data_ptr(), nseg, seg, dec.data_ptr(), d.numel(), header=True,
                                    trailer=False)
        t = eng.last_kernel_ms()["inflate"]
        ims = t if ims is None else min(ims, t)
    torch.cuda.synchronize()
    r = {"workload": name, "level": level, "strategy": STRATEGY_NAMES[strategy], "segment_bytes": seg, "mib": d.numel() >> 20,
         "ratio": round(clen / d.numel(), 4), "deflate_gbs": round(d.numel() / best / 1e6, 2),
         "inflate_gbs": round(d.numel() / ims / 1e6, 2), "round_trip_ok": bool(torch.equal(dec, d))}
    if parity_mib and not c.args.no_parity:
        r["parity_vs_oracle"] = parity_vs_oracle(d, out, ends, dict(seg=seg, level=level, strategy=strategy), 0, 2, parity_mib)["identical"]
    return r


def other_configs(c, text, mib=256):
    """The remaining BASELINE.json configurations at 256 MiB each on one GPU (synthetic data of each named shape)."""
    n = mib << 20
    res = [run_config(c, "text-like", text[:n].contiguous(), 6, 0, 65536, parity_mib=32)]
    mixed = device_input("mixed", n, 0, c.dev)
    res.append(run_config(c, "mixed-entropy", mixed, 6, 0, 262144, parity_mib=32))
    res.append(run_config(c, "mixed-entropy", mixed, 9, 0, 262144, parity_mib=8))
    del mixed
    png = device_input("png", n, 0, c.dev)
    res.append(run_config(c, "png-like Filtered", png, 6, 1, 65536, parity_mib=32))
    res.append(run_config(c, "png-like HuffmanOnly", png, 6, 2, 65536, parity_mib=32))
    res.append(run_config(c, "png-like Rle", png, 6, 3, 65536, parity_mib=32))
    return res


def bench_inflate(c):
    """Config 4: one full-flush-segmented stream, its segment index cut into contiguous compressed-offset ranges (one per rank),
    every rank decodes its range, Adler-32 folded with adler32_combine against the trailer (dist.inflate_sharded).  The stream
    is this library's level-1 output for text-like data -- bit-identical to the reference's (tests) -- built per rank from 1 GiB
    of input and repeated (full-flush segments are independent, so the concatenation is a valid segmented body)."""
    import torch
    from zlibstream_b200 import adler32_combine
    from zlibstream_b200.dist import inflate_sharded
    args, cfg, eng, dev, stream = c.args, c.cfg, c.eng, c.dev, c.stream
    seg = cfg["seg"]
    mib = cfg["mib"] if (c.world > 1 or c.args.size_mib != CONFIGS[4]["mib"]) else 16384   # N = 1: the whole 16 GiB stream on the one GPU
    base_mib = min(1024, mib)
    reps = max(1, mib // base_mib)
    nb = base_mib << 20
    base = device_input(cfg["gen"], nb, c.rank, dev)
    cap = eng.deflate_bound(nb, seg)
    comp1 = torch.empty(cap, dtype=torch.uint8, device=dev)
    ends1 = torch.empty(nb // seg, dtype=torch.int64, device=dev)
    clen, adler1 = eng.deflate_segments_device(base.data_ptr(), nb, seg, cfg["level"], cfg["strategy"], HASH, 1, comp1.data_ptr(), cap,
                                               ends1.data_ptr(), header=False, trailer=False)
    body = comp1[:clen]
    comp = body.repeat(reps)
    ends = torch.cat([ends1 + r * clen for r in range(reps)]).contiguous()
    n = nb * reps
    nseg = n // seg
    del comp1
    # the stream's trailer: fold of every rank's shard (each shard = reps copies of its 1 GiB input)
    shard_adler = 1
    for _ in range(reps):
        shard_adler = adler32_combine(shard_adler, adler1, nb)
    expect = None
    if c.world > 1:
        import torch.distributed as dist
        t = torch.tensor([shard_adler], dtype=torch.int64, device=dev)
        allt = [torch.empty_like(t) for _ in range(c.world)]
        dist.all_gather(allt, t)
        expect = 1
        for a in allt:
            expect = adler32_combine(expect, int(a.item()), n)
    else:
        expect = shard_adler
    dec = torch.empty(n, dtype=torch.uint8, device=dev)
    gather = None
    if args.gather and c.world > 1:
        gather = torch.empty(n * c.world, dtype=torch.uint8, device=dev) if c.rank == 0 else torch.empty(1, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def step():
        with torch.cuda.stream(stream):
            return inflate_sharded(eng, adler32_combine, comp, ends, seg, dec, has_header=False, expect_adler=expect, gather_to=gather)

    for _ in range(max(args.warmup, 0)):
        res = step()
    c.barrier()
    launches0 = eng.launch_count
    sampler = ClockSampler(c.local_rank)
    if c.rank == 0:
        sampler.start()
    kms = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
    for _ in range(args.steps):
        res = step()
        kms.append(eng.last_kernel_ms()["inflate"])
    with torch.cuda.stream(stream):
        ev1.record(stream)
    c.barrier()
    elapsed_ms = max_over_ranks(c, ev0.elapsed_time(ev1))
    clocks = sampler.stop() if c.rank == 0 else None
    launches = eng.launch_count - launches0
    ok = bool(torch.equal(dec[:nb], base)) and bool(torch.equal(dec[-nb:], base)) and res.adler == expect
    value = n * c.world * args.steps / (elapsed_ms / 1e3) / 1e9
    ratio = clen / nb

    e2e = None
    if not args.no_e2e and c.world == 1:
        import ctypes
        from zlibstream_b200 import _native
        en = min(n, 4096 << 20)      # bounded: 4 GiB of pinned output
        ek = en // seg
        elen = int(ends[ek - 1].item())
        h_in = torch.empty(elen, dtype=torch.uint8).pin_memory()
        h_in.copy_(comp[:elen])
        h_ends = ends[:ek].cpu().pin_memory()
        h_out = torch.empty(en, dtype=torch.uint8).pin_memory()
        ol, ad = ctypes.c_size_t(0), ctypes.c_uint32(0)
        flags = _native.ZSB_NO_HEADER | _native.ZSB_NO_TRAILER

        def call():
            rc = eng._lib.zsb_inflate_segments(eng._h, h_in.data_ptr(), elen, h_ends.data_ptr(), ek, seg, flags, h_out.data_ptr(), en,
                                               ctypes.byref(ol), ctypes.byref(ad))
            if rc not in (0, 1):
                raise RuntimeError("e2e inflate failed: " + eng.message)

        call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            call()
        dt = (time.perf_counter() - t0) / 2
        e2e = {"value": en / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": elen + 8 * ek, "d2h_bytes_per_step": en,
               "api": "zsb_inflate_segments with pinned host buffers (first 4 GiB of the stream; wall clock around the blocking call)"}

    if c.rank != 0:
        return None
    peak, peak_src = _hbm_peak()
    km = sum(kms) / len(kms)
    algo = n * (1 + ratio)
    achieved = algo / (km / 1e3) / 1e9
    line = {"metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": c.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": dict(workload_config(args, c.world), bytes_per_gpu=n, segments_per_gpu=nseg,
                                                               stream=f"{base_mib} MiB of input per rank compressed at level 1, body repeated {reps}x",
                                                               gathered_to_rank0=bool(gather is not None)),
            "compression_ratio": ratio, "gpu_launches": launches, "kernel_ms": {"inflate_kernel": km},
            "roofline": {"bound": "hbm", "kernel": "inflate_lanes_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": _traffic("inflate_config4", n),
                         "note": f"algorithmic bytes = output * (1 + ratio) = {algo:.0f} per launch; peak = {peak_src}"},
            "clocks": clocks, "round_trip_ok": ok, "adler_fold_matches_trailer": res.adler == expect}
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        smib = 256
        k = (smib << 20) // seg
        e = ends[:k].cpu().tolist()
        hb = comp[:e[-1]].cpu().numpy().tobytes()
        one = cpu_inflate(hb, e, 0, seg, smib << 20, 1, 1, 4.0)
        allc = cpu_inflate(hb, e, 0, seg, smib << 20, 1, cores, 8.0)
        line["cpu_baseline"] = {"value": allc["value"], "unit": UNIT, "cores": cores, "kind": "system zlib " + __import__("zlib").ZLIB_RUNTIME_VERSION,
                                "per_core_mbs": allc["per_core_mbs"], "one_core_sequential_mbs": one["value"] * 1e3,
                                "sample": f"first {smib} MiB of rank 0's stream, {allc['passes']} passes, native pinned threads, one segment "
                                          "at a time per thread"}
    return line


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
