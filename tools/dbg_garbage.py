import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'morphsym-hgnn_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
from helpers import *
name, B, layers = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = CONFIGS[name]
batch = make_batch(cfg, B, seed=3)
om = oracle_model(cfg, layers=layers, seed=1)
with torch.no_grad():
    out_o = om({k: v.double() for k, v in batch.x_dict.items()}, batch.edge_index_dict)
for mode in ['fp32', 'tc']:
    for fill in [0, 255, 0x3C, 0x7B]:
        nm = build_model(cfg, layers=layers, seed=2)
        nm.load_state_dict({k: v.float() for k, v in om.state_dict().items()})
        nm.set_mode(mode); nm = nm.to('cuda:0')
        b = batch.to('cuda:0')
        with torch.no_grad():
            nm(b.x_dict, b.edge_index_dict)              # compile engine, allocate ws
            eng = nm._last_engine
            for train in (False, True):
                eng.workspace(B, train, torch.device('cuda:0')).fill_(fill)
                nm._ensure_flat(torch.device('cuda:0'))
                out = eng.forward([b.x_dict[t] for t in nm.node_types], nm._flat, train=train)
                torch.cuda.synchronize()
                print(mode, 'fill', fill, 'train', train, 'err', f'{rel_err(out, out_o):.3e}', 'nan', torch.isnan(out).any().item(), flush=True)
