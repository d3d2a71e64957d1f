import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regtr_b200.config import get_config
from regtr_b200.regtr import GraphedRegTR, RegTR
from regtr_b200.synthetic import make_3dmatch_pair
from regtr_b200.weights import random_state_dict
DEV = 'cuda:0'
cfg = get_config('3dmatch')
model = RegTR(cfg).to(DEV).eval(); model.load_state_dict(random_state_dict(cfg, 2024), strict=True)
r = GraphedRegTR(model)
p = make_3dmatch_pair(2000)
b = {'src_xyz': [torch.from_numpy(p['src_xyz']).to(DEV)], 'tgt_xyz': [torch.from_numpy(p['tgt_xyz']).to(DEV)]}
for _ in range(3): r(dict(b))
torch.cuda.synchronize()
ts, tr = [], []
for _ in range(50):
    t0 = time.perf_counter(); tk = r.submit(dict(b)); t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter(); out = r.result(tk); t3 = time.perf_counter()
    ts.append(t1 - t0); tr.append(t3 - t2)
import statistics
print('submit ms', 1e3 * statistics.median(ts), 'result ms', 1e3 * statistics.median(tr))
st = list(r.graphs.values())[0]
t0 = time.perf_counter()
for _ in range(50): st['graph'].replay()
t1 = time.perf_counter(); torch.cuda.synchronize()
print('graph.replay() host ms', 1e3 * (t1 - t0) / 50)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(30):
    tk = r.submit(dict(b)); torch.cuda.synchronize(); r.result(tk)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
