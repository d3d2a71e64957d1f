"""Pre-processor alone on a real pair (target of compute-sanitizer runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from regtr_b200.config import get_config
from regtr_b200.kpconv import PreprocessorGPU
name = sys.argv[1] if len(sys.argv) > 1 else 'real_3dmatch_sun3d_home_38_41'
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'real', name + '_input.npz'))
cfg = get_config('modelnet' if 'modelnet' in name else '3dmatch')
pre = PreprocessorGPU(cfg)
for rep in range(2):
    meta = pre([torch.from_numpy(d['src_xyz']).cuda(), torch.from_numpy(d['tgt_xyz']).cuda()], lazy_upsamples=True)
    torch.cuda.synchronize()
    print(rep, [int(x.sum()) for x in meta['stack_lengths']])
