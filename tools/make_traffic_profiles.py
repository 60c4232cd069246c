"""profiles/<prefix>_step_traffic.json + <prefix>_dominant_traffic.json from an ncu launch list of one train step
(tools/launch_traffic.py does the aggregation).  usage: make_traffic_profiles.py launches.csv prefix "source note" """
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csv_path, prefix, note = sys.argv[1], sys.argv[2], sys.argv[3]
step_path = os.path.join(ROOT, "profiles", prefix + "_step_traffic.json")
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_traffic.py"), csv_path, step_path], check=True, stdout=subprocess.DEVNULL)
step = json.load(open(step_path))
KINDS = {"encoder_fwd": ("k_tc_encoder_stream", "k_tc_encoder_pair", "k_tc_encoder"), "stack_fwd": ("k_tc_stack2<0>", "k_tc_stack<0>"),
         "stack_bwd": ("k_tc_stack2<1>", "k_tc_stack<1>"), "dw_layers": ("k_tc_reducegemm",), "dw_encoder": ("k_tc_encoder_dw",)}
out = {}
for kind, names in KINDS.items():
    for name, a in step["kernels"].items():
        base = name.replace("void ", "")
        if any(base == n or base.startswith(n + "<") for n in names) and a["launches"]:
            out[kind] = {"dram_bytes_per_launch": (a["dram_read_bytes"] + a["dram_write_bytes"]) / a["launches"],
                         "dram_read_bytes": a["dram_read_bytes"] / a["launches"], "dram_write_bytes": a["dram_write_bytes"] / a["launches"],
                         "us_under_ncu": a["us"] / a["launches"], "kernel": base}
            break
out["step"] = {"dram_bytes": step["step"]["dram_read_bytes"] + step["step"]["dram_write_bytes"], "launches": step["step"]["launches"]}
out["source"] = note
json.dump(out, open(os.path.join(ROOT, "profiles", prefix + "_dominant_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "source"}, indent=1)[:1500])
