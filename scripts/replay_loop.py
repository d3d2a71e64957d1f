"""A few graph replays of the default bench workload (for ncu launch lists / profiles)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regtr_b200.config import get_config
from regtr_b200.regtr import GraphedRegTR, RegTR
from regtr_b200.synthetic import make_3dmatch_pair
from regtr_b200.weights import random_state_dict
DEV = 'cuda:0'
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = get_config('3dmatch')
model = RegTR(cfg).to(DEV).eval(); model.load_state_dict(random_state_dict(cfg, 2024), strict=True)
runner = GraphedRegTR(model)
p = make_3dmatch_pair(2000)
b = {'src_xyz': [torch.from_numpy(p['src_xyz']).to(DEV)], 'tgt_xyz': [torch.from_numpy(p['tgt_xyz']).to(DEV)]}
for _ in range(n_rep):
    out = runner(dict(b))
torch.cuda.synchronize()
print('ok', out['pose'][-1, 0].tolist())
