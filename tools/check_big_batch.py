import sys, os
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
for p in (os.path.join(ROOT, "morphsym-hgnn_b200"), os.path.join(ROOT, "tests")): sys.path.insert(0, p)
import torch
from ms_hgnn import _native as N
from ms_hgnn.synthetic import CONFIGS, build_model, make_batch
cfg = CONFIGS["mini_cheetah-k4-contact"]
for B in (70001, 33):
    b = make_batch(cfg, B, seed=1).to("cuda:0")
    nm = build_model(cfg, layers=8, seed=3).set_mode("tc").to("cuda:0")
    nm.validate_edges = "cached"
    outs = []
    for enc in (2, 0):
        N.set_option("encoder", enc)
        with torch.no_grad():
            outs.append(nm(b.x_dict, b.edge_index_dict).clone())
        torch.cuda.synchronize()
    print(B, "stream == pair:", torch.equal(outs[0], outs[1]), float(outs[0].abs().mean()))
    N.set_option("encoder", -1)
