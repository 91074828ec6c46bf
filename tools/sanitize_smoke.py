"""One small pass through every kernel of libzsb200, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Results are still compared with the oracle (test infrastructure), so a sanitizer run is also a parity run."""
import os
import sys
import zlib

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import corpus_data  # noqa: E402
import oracle_lib  # noqa: E402
from zlibstream_b200 import Engine  # noqa: E402


def main():
    orc = oracle_lib.load()
    eng = Engine(0)
    seg = 65536
    text = corpus_data.text_like(5 * seg + 4321, seed=11)
    mixed = corpus_data.mixed_entropy(2 * 262144 + 999, seed=12)
    png = corpus_data.png_like(4 * seg, seed=13)
    runs = b"\x00" * (seg + 77) + b"ab" * seg + bytes(range(256)) * 300
    cases = [("text", text, seg, 1, 0), ("text", text, seg, 3, 0), ("text", text, seg, 6, 0), ("text", text[:2 * seg + 99], seg, 9, 0),
             ("mixed", mixed, 262144, 1, 0), ("mixed", mixed, 262144, 6, 0), ("png", png, seg, 6, 1), ("png", png, seg, 6, 2),
             ("png", png, seg, 6, 3), ("runs", runs, seg, 6, 3), ("runs", runs, seg, 1, 4), ("text", text, seg, 1, 0, 1)]
    only = os.environ.get("ZSB_SMOKE_ONLY", "")   # "l1": the level-1 kernels and the inflate kernels only (a short racecheck)
    if only == "l1":
        cases = [c for c in cases if c[3] == 1]
    bad = 0
    for case in cases:
        name, data, ss, level, strategy = case[:5]
        hv = case[5] if len(case) > 5 else 0
        want, wends, wad = orc.deflate_segments(data, ss, level, strategy, hv)
        got, ends, ad = eng.deflate_segments(data, ss, level, strategy, hv)
        ok = got == want and ends == wends and ad == wad
        for kern in ("warp", "lanes", "lanes_small"):
            os.environ["ZSB_INFLATE_KERNEL"] = kern
            dec, dad = eng.inflate_segments(got, ends, ss, len(data))
            ok = ok and dec == data and dad == zlib.adler32(data)
        os.environ.pop("ZSB_INFLATE_KERNEL", None)
        dec, dad = eng.inflate_segments(got, None, ss, len(data))     # segment index from the markers
        ok = ok and dec == data
        print(f"{'ok  ' if ok else 'FAIL'} {name} seg={ss} L{level} S{strategy} hv={hv}", flush=True)
        bad += not ok
    if only:
        eng.close()
        print("sanitize_smoke:", "all identical" if not bad else f"{bad} FAILED")
        sys.exit(1 if bad else 0)
    # round-2 paths: level 0, sync / partial flush (chained segments, shared edge bytes), one-piece streams, the PNG front end,
    # the piped host-buffer inflate
    for name, kw, mode in (("sync", dict(sync_flush=True), 2), ("partial", dict(partial_flush=True), 1)):
        for level in (1, 6):
            want, wends = orc.deflate_writes(text, seg, level, 0, 0, flush_mode=mode)
            got, ends, _ = eng.deflate_segments(text, seg, level, 0, **kw)
            dec, _ = eng.inflate_stream(got, len(text))
            ok = got == want and ends == wends and dec == text
            print(f"{'ok  ' if ok else 'FAIL'} {name} flush L{level}", flush=True)
            bad += not ok
    got, ends, _ = eng.deflate_segments(text, seg, 0, 0)
    ok = got == orc.deflate_segments(text, seg, 0, 0)[0] and eng.inflate_segments(got, ends, seg, len(text))[0] == text
    for level in (1, 6):
        one, _ = eng.deflate_stream(text[:200001], level)
        ok = ok and one == orc.deflate_oneshot(text[:200001], level)[0] and eng.inflate_stream(one, 200001)[0] == text[:200001]
    print(f"{'ok  ' if ok else 'FAIL'} level 0, one-piece streams", flush=True)
    bad += not ok
    sys.path.insert(0, "oracle")
    import numpy as np
    import png_ref
    px = np.random.RandomState(5).randint(0, 256, size=(60, 500, 4)).astype(np.uint8).tobytes()
    png = eng.png_encode(px, 500, 60, 6, 5, level=6, strategy=1, idat_bytes=65536)
    ok = png_ref.decode(png)["pixels"] == px
    print(f"{'ok  ' if ok else 'FAIL'} png", flush=True)
    bad += not ok
    os.environ["ZSB_INFLATE_PIPE_SEGMENTS"] = "4"
    want, wends, _ = orc.deflate_segments(runs + text, seg, 6, 3)
    ok = eng.inflate_segments(want, wends, seg, len(runs + text))[0] == runs + text
    os.environ.pop("ZSB_INFLATE_PIPE_SEGMENTS")
    print(f"{'ok  ' if ok else 'FAIL'} piped host inflate (Rle carries)", flush=True)
    bad += not ok
    assert eng.adler32(text) == zlib.adler32(text)
    eng.close()
    print("sanitize_smoke:", "all identical" if not bad else f"{bad} FAILED")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
