"""Worker of tests/test_gpu_multirank.py (one process per GPU, launched with torch.distributed.run): the NCCL-reduced flat
gradient of the sharded batch must equal the single-rank gradient of the whole batch; overlapped and plain all-reduce agree
bit for bit; after an optimizer step every rank holds the same parameters."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "morphsym-hgnn_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ms_hgnn.synthetic import CONFIGS, build_model, make_batch  # noqa: E402
from ms_hgnn.train import FusedTrainer  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS["mini_cheetah-k4-contact"]
    B = 256 * world
    full = make_batch(cfg, B, seed=5)
    out = {"rank": rank, "world": world}
    for mode in ("tc", "fp32"):
        # rank r starts from DIFFERENT weights on purpose: the trainer must broadcast rank 0's before the first step
        nm = build_model(cfg, layers=8, seed=10 + rank).set_mode(mode).to(dev)
        nm.validate_edges = "cached"
        tr = FusedTrainer(nm, optimizer="sgd", lr=0.0)            # lr 0: the step leaves the weights alone, the gradients stay comparable
        shard = full.shard(rank, world).to(dev)
        tr.overlap_allreduce = True
        tr.train_step(shard)
        g_overlap = tr.grads.clone()
        tr.overlap_allreduce = False
        tr.train_step(shard)
        g_plain = tr.grads.clone()
        out[f"{mode}_overlap_equals_plain"] = bool(torch.equal(g_overlap, g_plain))
        # every rank now holds rank 0's weights
        w0 = nm.flat_parameters.clone()
        dist.broadcast(w0, src=0)
        out[f"{mode}_weights_synced"] = bool(torch.equal(w0, nm.flat_parameters))
        # the whole batch on one rank, no collective
        solo = FusedTrainer(nm, optimizer="sgd", lr=0.0, process_group=None)
        solo.world = 1
        solo._synced_params = True
        solo.train_step(full.to(dev))
        out[f"{mode}_err"] = rel(g_overlap, solo.grads)
        # graphed step == eager step
        tr.overlap_allreduce = True
        tr.train_step_graphed(shard)
        tr.train_step_graphed(shard)
        out[f"{mode}_graph_equals_eager"] = bool(torch.equal(tr.grads, g_overlap))
        # a real optimizer step keeps the replicas identical
        tr2 = FusedTrainer(nm, optimizer="adam", lr=1e-3)
        tr2.train_step(shard)
        w = nm.flat_parameters.clone()
        dist.broadcast(w, src=0)
        out[f"{mode}_replicas_identical_after_adam"] = bool(torch.equal(w, nm.flat_parameters))
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("DPRESULT " + json.dumps(gathered), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
