"""Runs a few native train steps of the bench workload (for ncu): python tools/profile_step.py [steps] [batch] [mode]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "morphsym-hgnn_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from ms_hgnn import _native as N  # noqa: E402
from ms_hgnn import morphology as M  # noqa: E402
from ms_hgnn.lightning_py.gnnLightning import HGNN_K4_Lightning  # noqa: E402
from ms_hgnn.synthetic import CONFIGS, make_batch  # noqa: E402
from ms_hgnn.train import FusedTrainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
mode = sys.argv[3] if len(sys.argv) > 3 else "tc"
cfg = CONFIGS["mini_cheetah-k4-contact"]
dev = torch.device("cuda", 0)
host = make_batch(cfg, B, seed=100)
module = HGNN_K4_Lightning(128, 8, M.K4_MINI_CHEETAH.metadata, host, "adam", 1e-4, regression=False, symmetry_mode="MorphSym",
                           group_operator_path=M.cfg_path(cfg.group)).to(dev)
module.model.validate_edges = "cached"
module.model.set_mode(mode)
trainer = FusedTrainer(module)
batch = host.to(dev)
n0 = N.launch_count()
for i in range(steps):
    loss = trainer.train_step(batch)
torch.cuda.synchronize()
print("loss", float(loss), "launches per step", (N.launch_count() - n0) / steps)
