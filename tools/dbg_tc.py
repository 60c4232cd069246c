import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'morphsym-hgnn_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
from helpers import *
from test_gpu_parity import native_run
name = sys.argv[1] if len(sys.argv) > 1 else 'mini_cheetah-k4-contact'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
layers = int(sys.argv[3]) if len(sys.argv) > 3 else 2
modes = sys.argv[4].split(',') if len(sys.argv) > 4 else ['fp32', 'tc', 'tc1x']
cfg = CONFIGS[name]
batch = make_batch(cfg, B, seed=3)
om = oracle_model(cfg, layers=layers, seed=1)
out_o, loss_o, g_o = oracle_run(cfg, om, batch)
for mode in modes:
    nm = build_model(cfg, layers=layers, seed=2)
    nm.load_state_dict({k: v.float() for k, v in om.state_dict().items()})
    nm.set_mode(mode)
    nm = nm.to('cuda:0')
    b = batch.to('cuda:0')
    with torch.no_grad():
        out = nm(b.x_dict, b.edge_index_dict)
    torch.cuda.synchronize()
    print(mode, 'inference out err', f'{rel_err(out, out_o):.3e}', flush=True)
    out_n, loss_n, g_n = native_run(cfg, nm, batch, mode=mode)
    torch.cuda.synchronize()
    errs = sorted(((rel_err(g_n[k], g_o[k]), k) for k in g_o if g_o[k].norm() > 0), reverse=True)
    print(mode, 'train out err', f'{rel_err(out_n, out_o):.3e}', 'loss', f'{abs(loss_n.item()-loss_o.item())/abs(loss_o.item()):.2e}',
          'grad worst', [(f'{e:.2e}', k) for e, k in errs[:3]], 'median', f'{errs[len(errs)//2][0]:.2e}', flush=True)
