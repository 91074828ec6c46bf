"""Level-1 two-kernel path diagnostics on a GPU box (development tool): diffs kernel A's candidate records against the
CPU model (tools/ab_model) and kernel B's symbols against the oracle trace.   python tools/l1_diag.py"""
import ctypes
import os
import struct
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
sys.path.insert(0, "tools")
import corpus_data  # noqa: E402
import oracle_lib  # noqa: E402
import gpu_diag  # noqa: E402
from zlibstream_b200 import Engine  # noqa: E402

ABM = ctypes.CDLL(os.path.join("tools", "ab_model", "libabmodel.so"))
ABM.abm_records.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]


def rec_str(r):
    d = struct.unpack("<10H", r[:20])
    hi = r[30] | (r[31] << 8)
    nk, more, usem = hi >> 11, (hi >> 10) & 1, hi & 0x3FF
    return " ".join(f"{x}{'*' if (usem >> k) & 1 else ''}:{r[20 + k] + 3 if (usem >> k) & 1 else 0}" for k, x in enumerate(d[:nk])) + \
        f" nk={nk} more={more}"


def check_records(eng, data, seg, nseg_check=3):
    """after a deflate call whose batch covered the whole input: records of the first segments vs the CPU model"""
    bad = 0
    for k in range(min(nseg_check, len(data) // seg)):
        chunk = data[k * seg:(k + 1) * seg]
        want = ctypes.create_string_buffer(32 * seg)
        ABM.abm_records(chunk, seg, 0, want)
        got = eng.debug_fetch(7, k * seg * 32, seg * 32)
        if got != want.raw:
            i = next(i for i in range(seg) if got[32 * i:32 * i + 32] != want.raw[32 * i:32 * i + 32])
            nbad = sum(got[32 * i:32 * i + 32] != want.raw[32 * i:32 * i + 32] for i in range(seg))
            print(f"   records: segment {k}: {nbad} positions differ, first at {i}:\n      got  {rec_str(got[32 * i:32 * i + 32])}\n      want {rec_str(want.raw[32 * i:32 * i + 32])}")
            bad += 1
    if not bad:
        print("   records of the first segments match the CPU model")


def main():
    orc = oracle_lib.load()
    eng = Engine(0)
    import test_gpu_deflate as T
    corpus = corpus_data.load_corpus()
    blob = b"".join(corpus[k] for k in sorted(corpus))
    cases = [("corpus_cat", blob), ("smallscratch", T._run_lengths_data(9, 8 * 65536) + b"\x05" * 70000 + corpus_data.text_like(3 * 65536 + 1234, seed=3)),
             ("runs1", T._run_lengths_data(1, 4 * 65536)),
             ("alice", corpus["alice29.txt"]),
             ("ptt5", corpus["ptt5"])]
    for name, data in cases:
        for seg in (65536, 262144):
            ok = gpu_diag.diff_case(eng, orc, name, data, seg, 1, 0)
            if not ok:
                check_records(eng, data, seg)
    eng.close()


if __name__ == "__main__":
    main()
