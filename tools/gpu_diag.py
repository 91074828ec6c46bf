"""Stage-by-stage parity diagnostics on a GPU box: diffs libzsb200's intermediate buffers against the oracle trace.

Usage (under gpurun):  python tools/gpu_diag.py [--quick]
Prints one line per case; on a mismatch prints the first differing stage (symbols / block records / block type /
layout / bytes) so that several bugs can be found per GPU round trip.
"""
import ctypes
import struct
import sys
import time
import zlib

sys.path.insert(0, ".")
sys.path.insert(0, "tests")

import corpus_data  # noqa: E402
import oracle_lib  # noqa: E402
from zlibstream_b200 import Engine  # noqa: E402

BLOCKREC = struct.Struct("<8I Q")   # sym_begin sym_count in_begin in_len buf_valid type hdr_bits body_bits bit_offset
assert BLOCKREC.size == 40


def diff_case(eng, orc, name, data, seg_size, level, strategy, hv=0, verbose=True):
    t0 = time.time()
    want, want_ends, segs = orc.trace_segments(data, seg_size, level, strategy, hv)
    t1 = time.time()
    try:
        got, ends, adler = eng.deflate_segments(data, seg_size, level, strategy, hv)
    except Exception as e:  # noqa: BLE001
        print(f"FAIL {name} seg={seg_size} L{level} S{strategy}: exception {e}")
        return False
    t2 = time.time()
    ok = got == want and ends == want_ends and adler == zlib.adler32(data)
    tag = "ok  " if ok else "FAIL"
    print(f"{tag} {name} n={len(data)} seg={seg_size} L{level} S{strategy} hv={hv}: out={len(got)} want={len(want)} "
          f"ratio={len(want) / max(1, len(data)):.3f} oracle={t1 - t0:.2f}s gpu={t2 - t1:.2f}s ms={eng.last_kernel_ms()}")
    if ok or not verbose:
        return ok
    if adler != zlib.adler32(data):
        print(f"   adler mismatch got {adler:08x} want {zlib.adler32(data):08x}")
    nseg = len(segs)
    max_blocks = seg_size // 16383 + 2
    nblocks = struct.unpack(f"<{nseg}I", eng.debug_fetch(2, 0, 4 * nseg))
    shown = 0
    for k in range(nseg):
        if shown >= 3:
            break
        wblocks = segs[k]
        # the oracle trace holds the real blocks followed by the Q10 extras (empty static blocks, n == 0 after the first
        # end-of-input flush); the device records only the parser's blocks
        recs = [BLOCKREC.unpack(eng.debug_fetch(1, (k * max_blocks + b) * 40, 40)) for b in range(min(nblocks[k], max_blocks))]
        sym_total = sum(r[1] for r in recs)
        wsyms = [s for b in wblocks for s in b["sym"]]
        raw = eng.debug_fetch(0, k * seg_size * 4, 4 * max(sym_total, 1))
        gs = struct.unpack(f"<{max(sym_total, 1)}I", raw)[:sym_total]
        gsyms = [(v & 0xFFFF, v >> 16) for v in gs]
        bad = False
        if gsyms != wsyms:
            bad = True
            i = next((i for i, (a, b) in enumerate(zip(gsyms, wsyms)) if a != b), min(len(gsyms), len(wsyms)))
            pos = sum((s[1] + 3) if s[0] else 1 for s in wsyms[:i])
            print(f"   seg {k}: symbol #{i} differs (input offset {pos}): got {gsyms[i:i + 4]} want {wsyms[i:i + 4]} "
                  f"(counts {len(gsyms)} vs {len(wsyms)})")
        # block records: compare with the oracle's first len(recs) blocks
        for b, r in enumerate(recs):
            if b >= len(wblocks):
                print(f"   seg {k}: device has extra block {b}: {r}")
                bad = True
                break
            w = wblocks[b]
            if (r[1], r[3], r[5]) != (w["n"], w["stored_len"], w["type"]):
                print(f"   seg {k} block {b}: got n={r[1]} stored_len={r[3]} type={r[5]} (hdr={r[6]} body={r[7]} valid={r[4]}) "
                      f"want n={w['n']} stored_len={w['stored_len']} type={w['type']}")
                bad = True
            seg_bit0 = (want_ends[k - 1] if k else 2) * 8
            if r[8] != w["bit_start"] - seg_bit0:
                print(f"   seg {k} block {b}: bit_offset got {r[8]} want {w['bit_start'] - seg_bit0}")
                bad = True
        lay = struct.unpack("<QIIQ8I", eng.debug_fetch(3, k * 56, 56))
        wlen = want_ends[k] - (want_ends[k - 1] if k else 2)
        if lay[1] != wlen:
            print(f"   seg {k}: out_bytes got {lay[1]} want {wlen}; oracle blocks {[(b['type'], b['n']) for b in wblocks]} "
                  f"tail_nbits={lay[2]} tail_off={lay[3]}")
            bad = True
        if not bad:
            a = want_ends[k - 1] if k else 2
            gseg, wseg = got[(ends[k - 1] if k else 2):ends[k]], want[a:want_ends[k]]
            if gseg != wseg:
                i = next((i for i, (x, y) in enumerate(zip(gseg, wseg)) if x != y), min(len(gseg), len(wseg)))
                print(f"   seg {k}: bytes differ at segment offset {i} of {len(wseg)}: got {gseg[i:i + 8].hex()} want {wseg[i:i + 8].hex()}; "
                      f"blocks {[(r[5], r[1], r[8]) for r in recs]}")
                bad = True
        if bad:
            shown += 1
    if got[:2] != want[:2]:
        print(f"   header got {got[:2].hex()} want {want[:2].hex()}")
    if got[-6:] != want[-6:]:
        print(f"   trailer got {got[-6:].hex()} want {want[-6:].hex()}")
    return ok


def main():
    quick = "--quick" in sys.argv
    orc = oracle_lib.load()
    eng = Engine(0)
    corpus = corpus_data.load_corpus()
    blob = b"".join(corpus[k] for k in sorted(corpus))
    print("libzsb200:", eng._lib.zsb_version().decode())

    # ---- Adler-32 ----
    import numpy as np
    rng = np.random.RandomState(1)
    for n in (0, 1, 8, 215, 1024, 1039, 2034, 4096, 5552, 5553, 65536, 65537, 1 << 20, (1 << 22) + 13):
        d = rng.bytes(n)
        for seed in (1, 0, 0x12345678 % (65521 << 16)):
            got = eng.adler32(d, seed) if n else seed
            want = zlib.adler32(d, seed)
            if got != want:
                print(f"FAIL adler n={n} seed={seed:#x}: got {got:08x} want {want:08x}")
    print("adler checks done")

    results = []
    text = corpus_data.text_like(1 << 20, 1)
    cases = [
        ("text1M", text, 65536, 1, 0),
        ("alice", corpus["alice29.txt"], 65536, 1, 0),
        ("blob", blob, 65536, 1, 0),
        ("blob", blob, 65536, 6, 0),
        ("blob", blob, 262144, 6, 0),
        ("blob", blob, 65536, 3, 0),
        ("blob", blob, 65536, 6, 1),
        ("blob", blob, 65536, 6, 2),
        ("blob", blob, 65536, 1, 4),
        ("blob", blob, 65536, 6, 3),
        ("kennedy", corpus["kennedy.xls"], 262144, 9, 0),
        ("random", rng.bytes(300000), 65536, 6, 0),
        ("zeros", bytes(400000), 65536, 6, 0),
        ("tiny", b"hello world", 65536, 6, 0),
        ("mixed", corpus_data.mixed_entropy(1 << 21), 262144, 6, 0),
        ("png", corpus_data.png_like(1 << 20), 65536, 6, 1),
        ("blob-mulhash", blob, 65536, 6, 0),
    ]
    if quick:
        cases = cases[:4]
    for c in cases:
        hv = 1 if c[0].endswith("mulhash") else 0
        results.append(diff_case(eng, orc, *c, hv=hv))
    print(f"deflate parity: {sum(results)}/{len(results)} cases ok")

    # ---- inflate ----
    nok = 0
    icases = [("blob", blob, 65536, 6, 0), ("text1M", text, 65536, 1, 0), ("random", rng.bytes(300000), 65536, 6, 0),
              ("blob256", blob, 262144, 9, 0), ("blobH", blob, 65536, 6, 2), ("blobF", blob, 65536, 1, 4)]
    for name, data, ss, lvl, st in icases:
        comp, ends, ad = orc.deflate_segments(data, ss, lvl, st)
        try:
            t0 = time.time()
            out, gad = eng.inflate_segments(comp, ends, ss, len(data))
            ok = out == data and gad == zlib.adler32(data)
            print(f"{'ok  ' if ok else 'FAIL'} inflate {name} seg={ss} L{lvl} S{st}: {len(out)} bytes, {time.time() - t0:.2f}s ms={eng.last_kernel_ms()}")
            if not ok:
                i = next((i for i, (x, y) in enumerate(zip(out, data)) if x != y), min(len(out), len(data)))
                print(f"   first difference at {i} (segment {i // ss}, offset {i % ss}); adler got {gad:08x}")
            nok += ok
        except Exception as e:  # noqa: BLE001
            print(f"FAIL inflate {name}: {e}")
    print(f"inflate parity: {nok}/{len(icases)} cases ok")
    eng.close()


if __name__ == "__main__":
    main()
