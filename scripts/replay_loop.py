"""A few graph replays of the bench workload (for ncu launch lists / profiles).

    python scripts/replay_loop.py [n_rep] [pairs_per_step] [config] [attention_impl] [profile]

With a 5th argument the LAST replay is bracketed by cudaProfilerStart/Stop: `ncu --profile-from-start off` then
captures exactly one CUDA-graph replay of the forward (capacity-shaped launches, real data).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from regtr_b200.config import get_config
from regtr_b200.regtr import GraphedRegTR, RegTR
from regtr_b200.synthetic import make_batch
from regtr_b200.weights import random_state_dict

DEV = 'cuda:0'
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
config = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = get_config('3dmatch')
if len(sys.argv) > 4 and sys.argv[4] != '-':
    cfg.attention_impl = sys.argv[4]
model = RegTR(cfg).to(DEV).eval()
model.load_state_dict(random_state_dict(cfg, 2024), strict=True)
runner = GraphedRegTR(model)
b = make_batch(config, B)
b = {k: [torch.from_numpy(c).to(DEV) for c in b[k]] for k in ('src_xyz', 'tgt_xyz')}
bracket = len(sys.argv) > 5
for i in range(n_rep):
    if bracket and i == n_rep - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    out = runner(dict(b))
torch.cuda.synchronize()
if bracket:
    torch.cuda.profiler.stop()
print('ok', out['pose'][-1, 0].tolist())
