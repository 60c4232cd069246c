"""Data-parallel host logic on CPU: 2 ranks over gloo.

The graph batch is sharded contiguously (HeteroBatch.shard), every rank produces a flat gradient buffer in
the library's canonical parameter order with its loss pre-scaled by 1/world, and ONE summed all-reduce of
that buffer must reproduce the full-batch gradient.  The per-rank gradients come from the oracle here (no GPU
in this container); on the GPU box the same path is exercised by bench.py --gpus N.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _flat_grads(cfg, model, batch, scale, layout):
    from helpers import oracle_loss
    x = {k: v.double() for k, v in batch.x_dict.items()}
    model.zero_grad()
    out = model(x, batch.edge_index_dict)
    loss = oracle_loss(cfg, out, batch.y.double(), batch.batch_size) * scale
    loss.backward()
    named = dict(model.named_parameters())
    n = sum(int(torch.Size(s).numel()) for _, _, s in layout)
    flat = torch.zeros(n, dtype=torch.float64)
    for name, off, shape in layout:
        g = named[name].grad
        if g is not None:
            flat[off:off + g.numel()] = g.reshape(-1)
    return flat, loss.detach()


def _worker(rank, world, port, ret):
    for p in (os.path.join(ROOT, "morphsym-hgnn_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import oracle_model
    from ms_hgnn import morphology as M
    from ms_hgnn.engine import Engine, build_spec
    from ms_hgnn.synthetic import CONFIGS, make_batch
    from ms_hgnn.train import allreduce_flat_
    torch.set_num_threads(2)
    cfg = CONFIGS["mini_cheetah-k4-contact"]
    tpl = M.K4_MINI_CHEETAH
    eng = Engine(build_spec(tpl.node_types, tpl.nodes_per_graph, cfg.in_width, tpl.edge_types, tpl.edges, ("gt", "gs"), 128, 2,
                            True, "base", "foot", 2, {}, None))      # host-only: plan + canonical flat layout
    layout = eng.param_layout()
    full = make_batch(cfg, 16, seed=7)
    om = oracle_model(cfg, layers=2, seed=1)
    shard = full.shard(rank, world)
    assert shard.batch_size == 8
    # bit-exact batching: the shard's edge_index is the template tiled over the shard
    for et, ei in shard.edge_index_dict.items():
        assert torch.equal(ei, tpl.edge_index(et, shard.batch_size))
    g, _ = _flat_grads(cfg, om, shard, 1.0 / world, layout)
    allreduce_flat_(g, world)
    if rank == 0:
        ref, _ = _flat_grads(cfg, om, full, 1.0, layout)
        ret["err"] = ((g - ref).norm() / ref.norm()).item()
        ret["n"] = g.numel()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_gradient_allreduce_equals_full_batch():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["n"] == 2 * (900 * 128 + 128) + 300 * 128 + 128 + 2 * 7 * (2 * 128 * 128 + 128) + 2 * (128 * 128 + 128) + 2 * 128 + 2
    assert ret["err"] < 1e-12
