"""profiles/r02_traffic.json from ncu launch lists (csv, --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum):
   python tools/traffic_from_ncu.py deflate_config2=gpurun_out/x.csv:1073741824:lz77_lists,lz77_resolve inflate_config4=...:bytes:inflate_lanes
For every key: dram bytes per launch of the named kernels (mean over the captured launches, summed over the kernels that make
up the stage) divided by the uncompressed bytes one launch processes."""
import csv
import json
import os
import sys

out = {}
for arg in sys.argv[1:]:
    key, rest = arg.split("=", 1)
    path, nbytes, kernels = rest.split(":")
    nbytes = int(nbytes)
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    per = {}
    for r in rows:
        name, metric, val = r[4], r[-3], float(r[-1].replace(",", ""))
        for k in kernels.split(","):
            if k in name:
                per.setdefault(k, {}).setdefault(metric, []).append(val)
    total, detail = 0.0, {}
    for k, m in per.items():
        rd = sum(m.get("dram__bytes_read.sum", [0])) / max(1, len(m.get("dram__bytes_read.sum", [0])))
        wr = sum(m.get("dram__bytes_write.sum", [0])) / max(1, len(m.get("dram__bytes_write.sum", [0])))
        t = sum(m.get("gpu__time_duration.sum", [0])) / max(1, len(m.get("gpu__time_duration.sum", [0])))
        detail[k] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "ns_under_ncu": t, "launches": len(m.get("gpu__time_duration.sum", []))}
        total += rd + wr
    out[key] = {"dram_bytes_per_launch": total, "unit_bytes_per_launch": nbytes, "dram_bytes_per_unit_byte": total / nbytes,
                "kernels": detail, "source": os.path.basename(path)}
print(json.dumps(out, indent=1))
