"""bench.py -- segmented DEFLATE throughput on B200 (BASELINE.json config 2) through libzsb200's C ABI.

  python bench.py --gpus 1 --steps 5 --warmup 3                 # this repo's CUDA path
  python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 # the reference's algorithm on the host cores (oracle)
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # one rank per GPU, segments sharded (weak scaling)

A step = one pass of the hot path over one batch: 1 GiB of synthetic text-like data per GPU, level 1 (deflate_fast),
64 KiB full-flush segments, Default strategy, CRC32C hash.  `value` = input GB/s with the input resident in HBM;
`e2e` = the same call with pinned HOST buffers (H2D of the input and D2H of the stream inside the timed region).
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries the ONE JSON line: keep NCCL's version banner (printed to stdout at NCCL_DEBUG=VERSION / WARN) out of it
# (VERSION and WARN both print it; an explicit INFO / TRACE of the caller is respected)
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    del os.environ["NCCL_DEBUG"]

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

SEG_SIZE = 65536
LEVEL = 1
STRATEGY = 0
HASH = 0
METRIC = "segmented_deflate_throughput"
UNIT = "GB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-mib", type=int, default=1024, help="input MiB per GPU and step")
    ap.add_argument("--cpu-sample-mib", type=int, default=0, help="0 = sized for ~10 s of CPU work")
    ap.add_argument("--no-inflate", action="store_true")
    ap.add_argument("--no-inflate-large", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": "text-like synthetic, 1 GiB per GPU, level 1 (deflate_fast), 64 KiB full-flush segments",
            "bytes_per_gpu": args.size_mib << 20, "segment_bytes": SEG_SIZE, "level": LEVEL, "strategy": "Default",
            "hash": "crc32c", "segments_per_gpu": (args.size_mib << 20) // SEG_SIZE, "sharding": f"segments/{world} ranks",
            "l2_policy": "input (1 GiB) and scratch are larger than the 126 MB L2"}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a C port of the reference's managed algorithm) on the host cores
# ---------------------------------------------------------------------------------------------------------------

def cpu_deflate_throughput(sample: bytes, threads: int):
    """GB/s of the oracle over `sample`, split into per-thread contiguous segment ranges (each its own stream)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from concurrent.futures import ThreadPoolExecutor
    orc = oracle_lib.load()
    nseg = len(sample) // SEG_SIZE
    per = max(1, (nseg + threads - 1) // threads)
    parts = [sample[i * per * SEG_SIZE:(i + 1) * per * SEG_SIZE] for i in range(threads)]
    parts = [p for p in parts if p]

    def work(p):
        out, _, _ = orc.deflate_segments(p, SEG_SIZE, LEVEL, STRATEGY, HASH)   # ctypes releases the GIL
        return len(out)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=len(parts)) as ex:
        outs = list(ex.map(work, parts))
    dt = time.perf_counter() - t0
    return len(sample) / dt / 1e9, dt, sum(outs)


def cpu_measure(sample: bytes, threads: int, target_s: float = 12.0, max_reps: int = 256):
    """Times the oracle on `sample` repeatedly until about target_s of CPU work has been done (first pass = calibration,
    counted).  Returns (GB/s, seconds, passes, compressed bytes of one pass)."""
    gbs, dt, comp = cpu_deflate_throughput(sample, threads)
    total_t, passes = dt, 1
    reps = int(min(max_reps, max(0, round(target_s / max(dt, 1e-3)) - 1)))
    for _ in range(reps):
        _, dt, comp = cpu_deflate_throughput(sample, threads)
        total_t += dt
        passes += 1
    return len(sample) * passes / total_t / 1e9, total_t, passes, comp


def host_sample(nbytes):
    import synth
    return synth.text_like(nbytes, seed=1, device="cpu", chunk=16 << 20).numpy().tobytes()


def run_reference(args):
    """--impl reference: the reference's own (CPU) algorithm for the same metric and config.  The managed reference
    cannot run here (no .NET); its C restatement in oracle/ is timed with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_mib = args.cpu_sample_mib or max(64, min(args.size_mib, 16 * cores))
    sample = host_sample(sample_mib << 20)
    for _ in range(args.warmup):
        cpu_deflate_throughput(sample[: (8 << 20)], cores)
    # every step = passes over the sample worth ~20 s / steps of CPU work, so the whole run stays within a minute or so
    t_total, n_total, comp, passes = 0.0, 0, 0, 0
    for _ in range(args.steps):
        gbs, dt, k, c = cpu_measure(sample, cores, target_s=max(2.0, 20.0 / max(1, args.steps)))
        t_total += dt
        n_total += len(sample) * k
        passes += k
        comp = c
    value = n_total / t_total / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{passes} passes over {sample_mib} MiB of the same text-like workload "
                                       f"({t_total:.1f} s of wall time in total), {cores} host threads, one contiguous "
                                       f"segment range (own stream) per thread, ratio {comp / len(sample):.3f}"},
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
# GPU arm
# ---------------------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    import synth
    from zlibstream_b200 import Engine, adler32_combine
    from zlibstream_b200.dist import deflate_sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: libzsb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.size_mib << 20
    nseg = n // SEG_SIZE
    eng = Engine(local_rank)
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)

    data = synth.text_like(n, seed=1 + rank, device=dev)            # synthetic, resident in HBM
    cap = eng.deflate_bound(n, SEG_SIZE)
    out = torch.empty(cap, dtype=torch.uint8, device=dev)
    seg_end = torch.empty(nseg, dtype=torch.int64, device=dev)
    gather = torch.empty(cap * world + 64, dtype=torch.uint8, device=dev) if (world > 1 and rank == 0) else \
        (torch.empty(1, dtype=torch.uint8, device=dev) if world > 1 else None)
    first_seg = rank * nseg
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        with torch.cuda.stream(stream):
            return deflate_sharded(eng, adler32_combine, data, SEG_SIZE, LEVEL, STRATEGY, HASH, first_seg, out, seg_end,
                                   gather_to=gather)

    launches0 = 0
    for _ in range(max(args.warmup, 0)):
        res = step()
    barrier()
    launches0 = eng.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lz_ms, hf_ms = [], []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
    for _ in range(args.steps):
        res = step()
        k = eng.last_kernel_ms()
        lz_ms.append(k["lz77"])
        hf_ms.append(k["huffman"])
    with torch.cuda.stream(stream):
        ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count - launches0
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    total_bytes = n * world * args.steps
    value = total_bytes / (elapsed_ms / 1e3) / 1e9
    ratio = res.local_len / n

    # ---- inflate of the stream just produced (bit-identical to the reference's, see tests) ----
    inflate = None
    if not args.no_inflate:
        comp_len = res.local_len
        dec = torch.empty(n, dtype=torch.uint8, device=dev)
        with torch.cuda.stream(stream):
            hdr = rank == 0
            for _ in range(2):
                eng.inflate_segments_device(out.data_ptr(), comp_len, seg_end.data_ptr(), nseg, SEG_SIZE, dec.data_ptr(), n,
                                            header=hdr, trailer=False)
            barrier()
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            i0.record(stream)
            isteps = max(1, min(args.steps, 3))
            ims = []
            for _ in range(isteps):
                eng.inflate_segments_device(out.data_ptr(), comp_len, seg_end.data_ptr(), nseg, SEG_SIZE, dec.data_ptr(), n,
                                            header=hdr, trailer=False)
                ims.append(eng.last_kernel_ms()["inflate"])
            i1.record(stream)
        barrier()
        ims_total = i0.elapsed_time(i1)
        if world > 1:
            t = torch.tensor([ims_total], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ims_total = float(t.item())
        ok = bool(torch.equal(dec, data))
        inflate = {"value": n * world * isteps / (ims_total / 1e3) / 1e9, "unit": UNIT, "ms_per_step": ims_total / isteps,
                   "kernel_ms": sum(ims) / len(ims), "round_trip_ok": ok, "steps": isteps,
                   "roofline_frac": (n * (1 + ratio) / (sum(ims) / len(ims) / 1e3) / 1e9) / _hbm_peak()[0]}

    # ---- inflate at the sizes of BASELINE config 4: the 16 GiB stream over 8 GPUs (2 GiB decoded per GPU) and, on one GPU,
    # the whole 16 GiB: the stream body above, 2 or 16 times in a row -- full-flush segments are independent, so the concatenation is a valid segmented body.
    # With 32768 segments the library switches to the lane-per-segment kernel.
    for reps in ((2, 16) if world == 1 else (2,)):
      if inflate is not None and not args.no_inflate_large and args.size_mib >= 1024:
          hdr_len = 2 if rank == 0 else 0
          body = out[hdr_len:res.local_len]
          big = body.repeat(reps)
          ends1 = seg_end - hdr_len
          big_ends = torch.cat([ends1 + r * body.numel() for r in range(reps)]).contiguous()
          big_dec = torch.empty(n * reps, dtype=torch.uint8, device=dev)
          with torch.cuda.stream(stream):
              bms = []
              for it in range(3):
                  eng.inflate_segments_device(big.data_ptr(), big.numel(), big_ends.data_ptr(), nseg * reps, SEG_SIZE,
                                              big_dec.data_ptr(), n * reps, header=False, trailer=False)
                  if it:
                      bms.append(eng.last_kernel_ms()["inflate"])
          torch.cuda.synchronize()
          okb = bool(torch.equal(big_dec[:n], data)) and bool(torch.equal(big_dec[-n:], data))
          bm = sum(bms) / len(bms)
          inflate["at_2gib_per_gpu" if reps == 2 else "at_16gib_one_gpu"] = {"value": n * reps / (bm / 1e3) / 1e9, "unit": UNIT, "kernel_ms": bm, "segments": nseg * reps,
                                        "round_trip_ok": okb, "scope": "this rank's kernel time (ranks are independent)",
                                        "roofline_frac": (n * reps * (1 + ratio) / (bm / 1e3) / 1e9) / _hbm_peak()[0]}
          del big, big_dec, big_ends

      # ---- the other BASELINE.json configurations, 256 MiB each, kernel time (context for the headline; not bench lines) ----
      other = None
      if world == 1 and not args.no_other_configs:
          other = other_configs(eng, data, dev)

    # ---- e2e: same call through the C ABI with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, eng, data, n, cap, world, rank, dev, stream, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = _hbm_peak()
    lz = sum(lz_ms) / len(lz_ms)
    algo_bytes = n * (1 + ratio)
    achieved = algo_bytes / (lz / 1e3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(args, world),
            "compression_ratio": ratio, "gpu_launches": launches,
            "kernel_ms": {"lz77_segments_kernel": lz, "huffman_stage": sum(hf_ms) / len(hf_ms)},
            "roofline": {"bound": "hbm", "kernel": "lz77_segments_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": _traffic_from_profile(n),
                         "note": f"algorithmic bytes = input * (1 + ratio) = {algo_bytes:.0f} per launch; peak = {peak_src}; "
                                 "the kernel is dependency-bound, not bandwidth-bound (DESIGN.md)"},
            "clocks": clocks}
    if inflate:
        line["inflate"] = inflate
    if e2e:
        line["e2e"] = e2e
    if other:
        line["other_configs"] = other
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        sample_mib = args.cpu_sample_mib or max(64, min(args.size_mib, 16 * cores))
        sample = data[: sample_mib << 20].cpu().numpy().tobytes()
        gbs, dt, k, c = cpu_measure(sample, cores, target_s=12.0)
        line["cpu_baseline"] = {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{k} passes over the first {sample_mib} MiB of rank 0's input ({dt:.1f} s), "
                                          f"{cores} host threads, one contiguous segment range (own stream) per thread; "
                                          "the C oracle restates the reference's managed algorithm (no .NET here)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, eng, data, n, cap, world, rank, dev, stream, barrier):
    import ctypes
    import torch
    import torch.distributed as dist
    from zlibstream_b200 import _native
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.copy_(data)
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    nseg = n // SEG_SIZE
    h_ends = torch.empty(nseg, dtype=torch.int64).pin_memory()
    lib = eng._lib
    out_len = ctypes.c_size_t(0)
    adler = ctypes.c_uint32(0)
    flags = (0 if rank == 0 else _native.ZSB_NO_HEADER) | _native.ZSB_NO_TRAILER

    def call():
        rc = lib.zsb_deflate_segments(eng._h, h_in.data_ptr(), n, SEG_SIZE, LEVEL, STRATEGY, HASH, rank * nseg, None, flags,
                                      h_out.data_ptr(), cap, ctypes.byref(out_len), h_ends.data_ptr(), ctypes.byref(adler))
        if rc != 0:
            raise RuntimeError("e2e deflate failed: " + eng.message)

    call()
    barrier()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            call()
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    wall = (time.perf_counter() - t0) * 1e3
    ms = max(ms, 0.0)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"value": n * world * steps / (ms / 1e3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": n,
            "d2h_bytes_per_step": int(out_len.value) + 8 * nseg + 16, "ms_per_step": ms / steps, "wall_ms_per_step": wall / steps,
            "api": "zsb_deflate_segments with pinned host buffers"}


def _hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "6.65 TB/s (of fallback, B200_PROFILING.md)"


def other_configs(eng, text, dev, mib=256):
    """Deflate + inflate kernel throughput of the remaining BASELINE.json configurations at 256 MiB each (synthetic data of
    each named shape: tests/corpus_data generators tiled on the device).  Every output is inflated back and compared."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import corpus_data
    n = mib << 20

    def tile(host_bytes):
        t = torch.frombuffer(bytearray(host_bytes), dtype=torch.uint8).to(dev)
        return t.repeat((n + t.numel() - 1) // t.numel())[:n].contiguous()

    def run(name, d, level, strategy, seg):
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
        return {"workload": name, "level": level, "strategy": strategy, "segment_bytes": seg, "mib": d.numel() >> 20,
                "ratio": round(clen / d.numel(), 4), "deflate_gbs": round(d.numel() / best / 1e6, 2),
                "inflate_gbs": round(d.numel() / ims / 1e6, 2), "round_trip_ok": bool(torch.equal(dec, d))}

    res = [run("text-like", text[:n].contiguous(), 6, 0, 65536)]
    mixed = tile(corpus_data.mixed_entropy(32 << 20, seed=2))
    res.append(run("mixed-entropy", mixed, 6, 0, 262144))
    res.append(run("mixed-entropy", mixed, 9, 0, 262144))
    del mixed
    png = tile(corpus_data.png_like(32 << 20, seed=3))
    res.append(run("png-like Filtered", png, 6, 1, 65536))
    res.append(run("png-like HuffmanOnly", png, 6, 2, 65536))
    res.append(run("png-like Rle", png, 6, 3, 65536))
    return res


def _traffic_from_profile(n):
    """dram bytes per launch of the LZ77 kernel: bytes per input byte from the committed ncu --set full capture
    (profiles/lz77_traffic.json, taken on a smaller launch of the same workload) times this launch's input bytes."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "lz77_traffic.json")))["dram_bytes_per_input_byte"] * n
    except Exception:
        return None


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
