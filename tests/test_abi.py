"""The C-ABI library loads without a GPU and exports every symbol include/zsb200.h declares; the host mirror keeps
the reference's enum values and argument errors.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "zsb200.h")).read()
    return sorted(set(re.findall(r"ZSB_API\s+[\w\s\*]+?\b(zsb_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    from zlibstream_b200 import _native
    lib = _native.load()
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert getattr(lib, n) is not None, n
    assert sorted(_native.EXPORTS) == names


def test_version_and_pure_host_entry_points():
    from zlibstream_b200 import _native
    lib = _native.load()
    assert lib.zsb_version().decode().startswith("zsb200 ") and "sm_100a" in lib.zsb_version().decode()
    assert lib.zsb_adler32_combine(1, 1, 0) == 1
    import zlib
    a, b = b"x" * 70001, bytes(range(256)) * 300
    assert lib.zsb_adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(a + b)
    assert lib.zsb_deflate_bound(1 << 20, 65536) > (1 << 20)


def test_enum_values_match_the_reference():
    from zlibstream_b200 import CompressionLevel, CompressionState, CompressionStrategy, FlushMode
    assert [int(CompressionLevel.DefaultCompression), int(CompressionLevel.NoCompression), int(CompressionLevel.BestSpeed),
            int(CompressionLevel.Level6), int(CompressionLevel.BestCompression)] == [-1, 0, 1, 6, 9]   # CompressionLevel.cs
    assert [int(s) for s in (CompressionStrategy.DefaultStrategy, CompressionStrategy.Filtered, CompressionStrategy.HuffmanOnly,
                             CompressionStrategy.Rle, CompressionStrategy.Fixed)] == [0, 1, 2, 3, 4]     # CompressionStrategy.cs
    assert [int(f) for f in (FlushMode.NoFlush, FlushMode.PartialFlush, FlushMode.SyncFlush, FlushMode.FullFlush,
                             FlushMode.Finish)] == [0, 1, 2, 3, 4]                                        # FlushMode.cs
    assert int(CompressionState.ZVERSIONERROR) == -6 and int(CompressionState.ZOK) == 0 and int(CompressionState.ZNEEDDICT) == 2


def test_no_cpu_fallback():
    """Without a device the engine refuses to start instead of silently computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from zlibstream_b200 import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(0)


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zlibstream_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                t = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle_lib|libzso|zso_|/oracle/|import oracle", t):
                    bad.append(f)
    assert not bad, bad


def test_bench_roofline_traffic_file_is_present():
    """bench.py's roofline.traffic comes from the round's committed ncu capture; a missing key would silently report null."""
    import json
    t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    for key in ("deflate_config2", "inflate_config4"):
        assert t[key]["dram_bytes_per_unit_byte"] > 0 and t[key]["unit_bytes_per_launch"] > 0
        assert all(k["launches"] >= 1 for k in t[key]["kernels"].values())
