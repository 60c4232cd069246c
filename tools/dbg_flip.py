import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'morphsym-hgnn_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
from helpers import *
from test_gpu_parity import native_run
name = sys.argv[1]
for B, seed in [(257, 257), (257, 1), (257, 2), (300, 3), (512, 4), (129, 5), (256, 6)]:
    cfg = CONFIGS[name]
    batch = make_batch(cfg, B, seed=seed)
    om = oracle_model(cfg, layers=8, seed=1)
    nm = build_model(cfg, layers=8, seed=2)
    nm.load_state_dict({k: v.float() for k, v in om.state_dict().items()})
    out_o, loss_o, g_o = oracle_run(cfg, om, batch)
    out_n, loss_n, g_n = native_run(cfg, nm, batch)
    errs = sorted(((rel_err(g_n[k], g_o[k]), k) for k in g_o if g_o[k].norm() > 0), reverse=True)
    print(B, seed, 'out', f'{rel_err(out_n, out_o):.2e}', 'worst', [(f'{e:.2e}', k) for e, k in errs[:3]], 'median', f'{errs[len(errs)//2][0]:.2e}', flush=True)
