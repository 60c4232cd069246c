"""Runs a few native inference forwards of the bench workload (for ncu): python tools/profile_infer.py [steps] [batch] [mode]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "morphsym-hgnn_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from ms_hgnn import _native as N  # noqa: E402
from ms_hgnn.synthetic import CONFIGS, build_model, make_batch  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
mode = sys.argv[3] if len(sys.argv) > 3 else "tc"
cfg = CONFIGS["mini_cheetah-k4-contact"]
nm = build_model(cfg, layers=8, seed=3).set_mode(mode).to("cuda:0")
nm.validate_edges = "cached"
b = make_batch(cfg, B, seed=1).to("cuda:0")
n0 = N.launch_count()
with torch.no_grad():
    for _ in range(steps):
        out = nm(b.x_dict, b.edge_index_dict)
torch.cuda.synchronize()
print("out", float(out.float().abs().mean()), "launches per forward", (N.launch_count() - n0) / steps)
