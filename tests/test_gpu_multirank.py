"""Hardware multi-rank correctness: NCCL over NVLink, one process per GPU.  Skipped with fewer than 2 GPUs; the
host-side logic of the same path runs on CPU over gloo in tests/test_dp_gloo.py."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_nccl_reduced_gradients_equal_the_full_batch_gradient():
    n = min(torch.cuda.device_count(), 4)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "_dp_nccl_worker.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    line = next(l for l in r.stdout.splitlines() if l.startswith("DPRESULT "))
    res = json.loads(line[len("DPRESULT "):])
    assert len(res) == n
    for out in res:
        for mode in ("tc", "fp32"):
            # flat gradient of the sharded, all-reduced step against one rank running the whole batch: the per-graph arithmetic
            # is identical, only the order of the sums over graphs differs
            assert out[f"{mode}_err"] <= 2e-5, out
            assert out[f"{mode}_overlap_equals_plain"], out
            assert out[f"{mode}_weights_synced"], out
            assert out[f"{mode}_graph_equals_eager"], out
            assert out[f"{mode}_replicas_identical_after_adam"], out
    print("multi-rank:", {k: v for k, v in res[0].items() if k.endswith("_err")})
