"""Where the single-thread roles of the stack kernel wait (MSHGNN_STACK_TIMING=1 build): python tools/stack_timing.py [B] [mode]."""
import os
import sys

os.environ["MSHGNN_STACK_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "morphsym-hgnn_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from ms_hgnn import _native as N  # noqa: E402
from ms_hgnn.synthetic import CONFIGS, build_model, make_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mode = sys.argv[2] if len(sys.argv) > 2 else "tc"
cfg = CONFIGS["mini_cheetah-k4-contact"]
dev = torch.device("cuda:0")
nm = build_model(cfg, layers=8, seed=3).set_mode(mode).to(dev)
nm.validate_edges = "cached"
b = make_batch(cfg, B, seed=1).to(dev)
NAMES = ["producer: ring slot", "scheduler: dependency", "MMA: operands", "MMA: accumulator free", "MMA: staged operand",
         "epilogue g0: accumulator", "kernel", "steps", "MMA: item queue", "MMA: issuing a K block", "producer: issuing TMA", "producer: item queue"]


def show(tag, train):
    eng = nm._last_engine
    off = eng.plan.stack_timing_offset(B, train, eng.mode)
    torch.cuda.synchronize()
    t = eng._ws[off:off + 148 * 128].view(torch.int64).view(148, 16).double().cpu()
    tot = t[:, 6].mean().item()
    if os.environ.get("STACK_TIMING_COMPACT"):
        short = {1: "sched:dep", 11: "prod:queue", 0: "prod:ring", 8: "mma:queue", 2: "mma:operands", 4: "mma:staged", 3: "mma:accfree", 9: "mma:issue", 5: "epi:acc"}
        print(f"{tag:18s} {tot / 1e6:.3f} Mcyc | " + "  ".join(f"{v} {100 * t[:, j].mean().item() / tot:4.1f}" for j, v in short.items()))
        return
    print(f"--- {tag}: kernel {tot / 1e6:.3f} Mcycles per CTA (mean), {t[:, 7].mean().item():.1f} steps per CTA")
    for j in (1, 11, 0, 10, 8, 2, 3, 4, 9, 5):
        print(f"   {NAMES[j]:28s} {100 * t[:, j].mean().item() / tot:5.1f} % of kernel cycles (min {100 * t[:, j].min().item() / tot:4.1f}, max {100 * t[:, j].max().item() / tot:4.1f})")


for it in range(2):
    with torch.no_grad():
        nm(b.x_dict, b.edge_index_dict)
    if it:
        show("inference forward", False)
    out = nm(b.x_dict, b.edge_index_dict)
    if it:
        show("training forward", True)
    eng = nm._last_engine
    loss, dout = eng.loss(out.detach().contiguous(), b.y, N.LOSS_CE2)
    out.backward(dout)
    if it:
        show("backward dX chain", True)
