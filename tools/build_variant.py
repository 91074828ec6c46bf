"""Builds an A/B variant of libzsb200.so: python tools/build_variant.py NAME file.cu -DX=1 [file2.cu -DY=2 ...]
Only the named sources are recompiled (with the macros that follow them); the other objects come from the regular build.
Load it with ZSB200_LIB=tools/variants/libzsb200_NAME.so."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zlibstream_b200", "csrc"))
import build as zb  # noqa: E402


def main():
    name = sys.argv[1]
    overrides = {}
    cur = None
    for a in sys.argv[2:]:
        if a.endswith(".cu"):
            cur = a
            overrides[cur] = []
        else:
            overrides[cur].append(a)
    zb.build()
    outdir = os.path.join(ROOT, "tools", "variants")
    os.makedirs(outdir, exist_ok=True)
    objs = []
    for src in zb.SOURCES:
        o = os.path.join(zb.HERE, "build", src.replace(".cu", ".o"))
        if src in overrides:
            o = os.path.join(outdir, f"{name}_{src.replace('.cu', '.o')}")
            subprocess.check_call([zb.NVCC] + zb.FLAGS + overrides[src] + ["-c", os.path.join(zb.HERE, src), "-o", o])
        objs.append(o)
    out = os.path.join(outdir, f"libzsb200_{name}.so")
    subprocess.check_call([zb.NVCC, "-shared", "--cudart=static", "-o", out] + objs)
    print(out)


if __name__ == "__main__":
    main()
