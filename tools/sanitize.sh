#!/bin/bash
# compute-sanitizer passes over a small train step (tensor-core mode), the window builder and the step metrics.
set -u
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
for p in (os.path.join(ROOT, "morphsym-hgnn_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
from ms_hgnn import _native as N, morphology as M
from ms_hgnn.lightning_py.gnnLightning import HGNN_K4_Lightning
from ms_hgnn.synthetic import CONFIGS, make_batch
from ms_hgnn.train import FusedTrainer
from ms_hgnn.windows import DeviceSequence, WindowSpec
import window_oracle as WO
cfg = CONFIGS["mini_cheetah-k4-contact"]
dev = torch.device("cuda", 0)
for B in (300, 700):
    host = make_batch(cfg, B, seed=1)
    mod = HGNN_K4_Lightning(128, 8, M.K4_MINI_CHEETAH.metadata, host, "adam", 1e-4, regression=False, symmetry_mode="MorphSym",
                            group_operator_path=M.cfg_path(cfg.group)).to(dev)
    mod.model.set_mode("tc")
    tr = FusedTrainer(mod)
    b = host.to(dev)
    for _ in range(2):
        loss = tr.train_step(b)
    with torch.no_grad():
        y, yp = mod.step_helper_function(b)
        mod.calculate_losses_step(y, yp)
    torch.cuda.synchronize()
    print("B", B, "loss", float(loss), "acc", float(mod.acc))
mat = WO.synthetic_mat(700, seed=2, dtype=np.float32)
ds = DeviceSequence(mat, WindowSpec("heterogeneous_gnn_k4", 150, True), dev, torch.float32)
wb = ds.batch(torch.arange(0, 500, 3))
torch.cuda.synchronize()
print("windows", float(wb.x_dict["foot"].abs().mean()))
PY
# one-CTA stack kernel (these batches are below the CTA-pair threshold), then the CTA-pair kernel forced (MSHGNN_STACK_2CTA=2)
for pair in ${SAN_PAIRS:-1 2}; do
for tool in ${SAN_TOOLS:-memcheck racecheck}; do
  MSHGNN_STACK_2CTA=$pair timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitize_${tool}_pair$pair.log 2>&1
  echo "== $tool (MSHGNN_STACK_2CTA=$pair): $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_pair$pair.log) summary lines"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|loss|windows" gpurun_out/sanitize_${tool}_pair$pair.log | head -12
done
done
