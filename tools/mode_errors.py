"""Per-tensor errors of a native arithmetic mode against the fp64 oracle on the parity cases (predictions, loss, worst gradient
tensor - plain, i.e. against the oracle's own ReLU pattern).  usage: python tools/mode_errors.py [mode] > profiles/..."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "morphsym-hgnn_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from helpers import oracle_model, oracle_run, rel_err  # noqa: E402
from ms_hgnn import _native as N  # noqa: E402
from ms_hgnn.synthetic import CONFIGS, build_model, make_batch  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tc1x"
CASES = [("mini_cheetah-k4-contact", 200, 8), ("mini_cheetah-k4-contact", 2048, 8), ("mini_cheetah-c2-contact", 96, 8), ("a1-c2-grf", 64, 8),
         ("k4-grf-regression", 130, 8), ("solo12-k4-com", 257, 8), ("mi-grf", 20, 8), ("mi-contact", 400, 8)]
rows = []
for name, B, L in CASES:
    cfg = CONFIGS[name]
    batch = make_batch(cfg, B, seed=B)
    om = oracle_model(cfg, layers=L, seed=1)
    out_o, loss_o, g_o = oracle_run(cfg, om, batch)
    nm = build_model(cfg, layers=L, seed=2)
    nm.load_state_dict({k: v.float() for k, v in om.state_dict().items()})
    nm = nm.set_mode(mode).to("cuda:0")
    b = batch.to("cuda:0")
    nm.zero_grad()
    out = nm(b.x_dict, b.edge_index_dict)
    eng = nm._last_engine
    C = eng.spec["out_channels"]
    loss, dout = eng.loss(out.detach().reshape(-1, C).float().contiguous(), b.y, N.LOSS_CE2 if cfg.loss == "ce" else N.LOSS_MSE)
    out.backward(dout.view_as(out).to(out.dtype))
    errs = {k: rel_err(p.grad, g_o[k]) for k, p in nm.named_parameters() if p.grad is not None and g_o[k].norm() > 0}
    worst = max(errs, key=errs.get)
    srt = sorted(errs.values())
    rows.append({"case": name, "B": B, "L": L, "mode": mode, "out_err": rel_err(out, out_o), "loss_err": abs(loss.item() - loss_o.item()) / abs(loss_o.item()),
                 "grad_err_worst": errs[worst], "grad_err_worst_tensor": worst, "grad_err_median": srt[len(srt) // 2],
                 "grad_tensors_over_1e-3": sum(1 for v in srt if v > 1e-3), "grad_tensors": len(srt)})
    print(json.dumps(rows[-1]), flush=True)
