"""Runs the window builder on the bench workload (for ncu): python tools/profile_windows.py [calls]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "morphsym-hgnn_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from ms_hgnn.windows import DeviceSequence, WindowSpec  # noqa: E402

n_rows, B = 1_000_000, 16384
rng = np.random.default_rng(7)
mat = {k: rng.standard_normal((n_rows, w), dtype=np.float32) for k, w in (("imu_acc", 3), ("imu_omega", 3), ("q", 12), ("qd", 12), ("p", 12), ("v", 12))}
mat["contacts"] = (rng.random((n_rows, 4)) < 0.5).astype(np.float32)
ds = DeviceSequence(mat, WindowSpec("heterogeneous_gnn_k4", 150, True), "cuda:0", torch.float32)
idx = torch.from_numpy(rng.integers(0, len(ds), size=B)).cuda()
b = ds.batch(idx)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ds.batch(idx, out=b)
torch.cuda.synchronize()
print("ok", float(b.x_dict["joint"].abs().mean()))
