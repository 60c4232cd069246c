"""Aggregates an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) into a
per-kernel table of one train step: launches, time, DRAM bytes, implied GB/s.  usage: launch_traffic.py launches.csv out.json"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
i0 = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[i0], rows[i0 + 1:]
kn, mn, mu, mv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
per = collections.OrderedDict()
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in data:
    if len(r) <= mv:
        continue
    key = (r[0], r[kn].split("(")[0])
    per.setdefault(key, {})[r[mn]] = float(r[mv].replace(",", "")) * scale[r[mu]]
# the capture window may cover a bit more than one step: keep exactly one period starting at the first encoder forward
ids = list(per)
starts = [i for i, k in enumerate(ids) if k[1].endswith("k_tc_encoder")]
lo, hi = (starts[0], starts[1]) if len(starts) > 1 else (0, len(ids))
agg = collections.OrderedDict()
for k in ids[lo:hi]:
    m = per[k]
    a = agg.setdefault(k[1], {"launches": 0, "us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
    a["launches"] += 1; a["us"] += m.get("gpu__time_duration.sum", 0.0)
    a["dram_read_bytes"] += m.get("dram__bytes_read.sum", 0.0); a["dram_write_bytes"] += m.get("dram__bytes_write.sum", 0.0)
tot = {"launches": 0, "us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0}
for a in agg.values():
    for f in tot:
        tot[f] += a[f]
    a["dram_gbs_under_ncu"] = (a["dram_read_bytes"] + a["dram_write_bytes"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else None
tot["dram_gbs_under_ncu"] = (tot["dram_read_bytes"] + tot["dram_write_bytes"]) / (tot["us"] * 1e-6) / 1e9
out = {"note": "one train step (16384 graphs, K4 Mini Cheetah, MODE_TC) under ncu --clock-control none: per-launch times are serialised and "
               "cold-cache, DRAM byte counts are the hardware counters", "kernels": agg, "step": tot}
json.dump(out, open(sys.argv[2], "w"), indent=1)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{k:38s} n={a['launches']:3d} {a['us']:9.1f} us  R {a['dram_read_bytes']/1e6:8.1f} MB  W {a['dram_write_bytes']/1e6:8.1f} MB  {a['dram_gbs_under_ncu']:7.0f} GB/s")
print("step", {k: round(v, 1) for k, v in tot.items()})
