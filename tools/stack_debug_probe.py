"""Runs two train steps under an MSHGNN_STACK_DEBUG ablation mask and reports whether a dependency wait timed out."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "morphsym-hgnn_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from ms_hgnn import _native as N
from ms_hgnn.synthetic import CONFIGS, build_model, make_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cfg = CONFIGS["mini_cheetah-k4-contact"]
nm = build_model(cfg, layers=8, seed=3).set_mode("tc").to("cuda:0")
nm.validate_edges = "cached"
b = make_batch(cfg, B, seed=1).to("cuda:0")
for it in range(2):
    t0 = time.time()
    out = nm(b.x_dict, b.edge_index_dict)
    torch.cuda.synchronize()
    eng = nm._last_engine
    st_f = eng.plan.stack_status(B, True, eng.mode, eng._ws.data_ptr())
    t1 = time.time()
    loss, dout = eng.loss(out.detach().contiguous(), b.y, N.LOSS_CE2)
    out.backward(dout)
    torch.cuda.synchronize()
    st_b = eng.plan.stack_status(B, True, eng.mode, eng._ws.data_ptr())
    print(f"debug={os.environ.get('MSHGNN_STACK_DEBUG')} step {it}: forward {t1 - t0:.3f}s status {st_f}, backward {time.time() - t1:.3f}s status {st_b}", flush=True)
