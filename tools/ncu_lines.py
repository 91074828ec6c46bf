"""Per-source-line view of an ncu capture (development tool): joins the SASS page of an .ncu-rep (instructions executed,
stall samples per instruction) with nvdisasm's line table of the same build of libzsb200.so.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep lz77_resolve [top [mangled-name regex]]      # kernel-name regex
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(kernel_re):
    """offset -> (file, line) for the first kernel of libzsb200.so whose mangled name matches."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "zlibstream_b200", "libzsb200.so")], cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        if "sm_100a" not in f:
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, table, loc = None, {}, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                if cur_fn and re.search(kernel_re, cur_fn) and table:
                    return table
                cur_fn, table, loc = m.group(1), {}, None
                continue
            m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
            if m:
                loc = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
            if m and cur_fn:
                table[int(m.group(1), 16)] = loc
        if cur_fn and re.search(kernel_re, cur_fn) and table:
            return table
    return {}


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    mre = sys.argv[4] if len(sys.argv) > 4 else kre     # regex on the MANGLED name (template instances: ...ILb0E...)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hi]
    ia, ie, ism, ith = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    table = line_table(mre)
    base = None
    agg = {}
    srcs = {}
    for r in rows[hi + 1:]:
        if len(r) <= ie or not r[ia].startswith("0x"):
            continue
        addr = int(r[ia], 16)
        if base is None:
            base = addr
        loc = table.get(addr - base) or ("?", 0)
        a = agg.setdefault(loc, [0, 0, 0, {}])
        a[0] += int(r[ie])
        a[1] += int(r[ism])
        a[2] += int(r[ith])
        for i, h in stall_cols:
            v = int(r[i] or 0)
            if v:
                a[3][h] = a[3].get(h, 0) + v
    tot_i = sum(a[0] for a in agg.values()) or 1
    tot_s = sum(a[1] for a in agg.values()) or 1
    print(f"{kre}: {tot_i} warp instructions, {tot_s} samples")
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        src = ""
        path = os.path.join(ROOT, "zlibstream_b200", "csrc", loc[0])
        if os.path.exists(path):
            if path not in srcs:
                srcs[path] = open(path).read().splitlines()
            if 0 < loc[1] <= len(srcs[path]):
                src = srcs[path][loc[1] - 1].strip()[:90]
        st = ", ".join(f"{k[6:]} {v * 100 // max(a[1], 1)}%" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
        print(f"{a[1] / tot_s * 100:5.1f}% smp {a[0] / tot_i * 100:5.1f}% inst thr/inst {a[2] / max(a[0], 1):4.1f}  {loc[0]}:{loc[1]:<4} {src}   [{st}]")


if __name__ == "__main__":
    main()
