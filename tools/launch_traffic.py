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
# the capture window covers a bit more than one step: keep exactly one period, from one optimizer launch to the next
ids = list(per)
ends = [i for i, k in enumerate(ids) if k[1].endswith("k_adam") or k[1].endswith("k_sgd")]
lo, hi = (ends[0] + 1, ends[1] + 1) if len(ends) > 1 else (0, len(ids))
agg = collections.OrderedDict()
# the persistent row-GEMM serves four launch kinds; inside one step they come in a fixed order (api.cu): forward
# [conv, base MLP] x L, backward [base MLP, dX] x L
n_pk = sum(1 for k in ids[lo:hi] if k[1].endswith("k_tc_rowgemm_persistent"))
pk_seen = 0
for k in ids[lo:hi]:
    m = per[k]
    name = k[1]
    if name.endswith("k_tc_rowgemm_persistent") and n_pk % 4 == 0:
        fwd = pk_seen < n_pk // 2
        name += ":" + (("conv_fwd", "base_mlp_fwd")[pk_seen % 2] if fwd else ("base_mlp_bwd", "dx_bwd")[pk_seen % 2])
        pk_seen += 1
    a = agg.setdefault(name, {"launches": 0, "us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
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
