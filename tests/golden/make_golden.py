"""Regenerates tests/golden/*: run HERE (needs /root/reference), commit the outputs.

  canterbury.tar.xz   the reference's tests/corpus, with alice29.txt / lcet10.txt / plrabn12.txt converted LF -> CRLF
                      to regain the canonical Canterbury sizes the published numbers were measured on
                      (/root/reference/.gitattributes:6 normalised the checkout)
  sizes.json          the reference's published `Bytes` column (benchmarks.md, SixLabors rows) -- the pin of the oracle
  segmented.json      oracle outputs (length + sha256 + segment end offsets) for full-flush-segmented runs, so that
                      the GPU parity tests can also be checked without rebuilding the oracle
"""
import hashlib
import io
import json
import os
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = "/root/reference/tests/corpus"
CRLF = {"alice29.txt", "lcet10.txt", "plrabn12.txt"}

# /root/reference/benchmarks.md lines 63-223 (corpus) and 18-28 (sparse image): file -> (canonical size, L1, L3, L6)
PUBLISHED = {
    "alice29.txt": (152089, 63340, 60207, 55818), "asyoulik.txt": (125179, 55139, 52914, 50068),
    "cp.html": (24603, 8907, 8645, 8233), "fields.c": (11150, 3766, 3570, 3280),
    "grammar.lsp": (3721, 1334, 1316, 1251), "kennedy.xls": (1029744, 199356, 203717, 187289),
    "lcet10.txt": (426754, 167403, 160050, 147916), "plrabn12.txt": (481861, 220181, 209933, 199026),
    "ptt5": (513216, 67013, 60164, 59946), "sum": (38240, 14694, 14383, 14002), "xargs.1": (4227, 1901, 1876, 1828),
    "sparse_rgba_3500": (49000000, 825050, 825050, 659280),
}


def main():
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:xz", preset=9) as tf:
        for name in sorted(os.listdir(REF)):
            data = open(os.path.join(REF, name), "rb").read()
            if name in CRLF:
                data = data.replace(b"\n", b"\r\n")
            assert len(data) == PUBLISHED[name][0], (name, len(data))
            ti = tarfile.TarInfo(name)
            ti.size = len(data)
            tf.addfile(ti, io.BytesIO(data))
    open(os.path.join(HERE, "canterbury.tar.xz"), "wb").write(buf.getvalue())
    json.dump({k: {"size": v[0], "L1": v[1], "L3": v[2], "L6": v[3]} for k, v in PUBLISHED.items()},
              open(os.path.join(HERE, "sizes.json"), "w"), indent=1)

    import corpus_data
    import oracle_lib
    orc = oracle_lib.load()
    corpus = corpus_data.load_corpus()
    seg = {}
    blob = b"".join(corpus[k] for k in sorted(corpus))  # 2.8 MB mixed file
    cases = [("corpus_cat", blob, 65536), ("corpus_cat", blob, 262144)]
    for name, data, ss in cases:
        for level, strat in [(1, 0), (6, 0), (9, 0), (6, 1), (6, 2), (6, 3), (1, 4), (3, 0)]:
            out, ends, ad = orc.deflate_segments(data, ss, level, strat)
            seg[f"{name}/seg{ss}/L{level}/S{strat}"] = {
                "len": len(out), "sha256": hashlib.sha256(out).hexdigest(), "adler": ad,
                "ends_sha256": hashlib.sha256(",".join(map(str, ends)).encode()).hexdigest(), "nseg": len(ends)}
    json.dump(seg, open(os.path.join(HERE, "segmented.json"), "w"), indent=1)
    print("wrote", len(buf.getvalue()), "bytes of corpus fixture and", len(seg), "segmented vectors")


if __name__ == "__main__":
    main()
