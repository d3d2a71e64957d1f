import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from regtr_b200.config import get_config
from regtr_b200.regtr import GraphedRegTR, RegTR
from regtr_b200.synthetic import make_3dmatch_pair
from regtr_b200.weights import random_state_dict
DEV = 'cuda:0'
cfg = get_config('3dmatch')
sd = random_state_dict(cfg, 8)
model = RegTR(cfg).to(DEV).eval(); model.load_state_dict(sd, strict=True)
runner = GraphedRegTR(model, bucket=16384)
G = lambda a: torch.from_numpy(a).to(DEV)
for i, n in enumerate((5000, 6000)):
    p = make_3dmatch_pair(2200 + i, n)
    b_e = {'src_xyz': [G(p['src_xyz'])], 'tgt_xyz': [G(p['tgt_xyz'])]}
    b_g = {'src_xyz': [G(p['src_xyz'])], 'tgt_xyz': [G(p['tgt_xyz'])]}
    want = model(b_e); got = runner(b_g)
    me, mg = b_e['kpconv_meta'], b_g['kpconv_meta']
    print('lens eager', me['_lens'], 'graph', mg['_lens'])
    for key in ('points', 'neighbors', 'pools', 'upsamples'):
        for l, (a, b) in enumerate(zip(mg[key], me[key])):
            same = a.shape == b.shape and torch.equal(a, b)
            extra = ''
            if not same and a.shape == b.shape:
                d = (a != b)
                extra = f'ndiff {int(d.sum())} of {d.numel()} rows {int(d.any(-1).sum())}'
                if key == 'points':
                    extra += f' maxabs {float((a-b).abs().max())}'
            print(i, key, l, tuple(a.shape), tuple(b.shape), same, extra)
    print('pose diff', float((got['pose'] - want['pose']).abs().max()))
# timing of replay
p = make_3dmatch_pair(2000)
b = {'src_xyz': [G(p['src_xyz'])], 'tgt_xyz': [G(p['tgt_xyz'])]}
r2 = GraphedRegTR(model)
for _ in range(3): r2(dict(b))
torch.cuda.synchronize()
st = list(r2.graphs.values())[0]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): st['graph'].replay()
e1.record(); torch.cuda.synchronize()
print('replay ms', e0.elapsed_time(e1) / 10, 'caps', st['caps'])
t = time.perf_counter()
for _ in range(10): r2(dict(b))
torch.cuda.synchronize(); print('call ms', (time.perf_counter() - t) * 100)
t = time.perf_counter()
for _ in range(10): model(dict(b))
torch.cuda.synchronize(); print('eager call ms', (time.perf_counter() - t) * 100)
