"""Where does the fp32 CUDA path lose accuracy?  (diagnostic, not collected by pytest)

    python tests/diag_accuracy.py [case]

Runs one golden case through the CUDA forward under a few A/B switches and prints, per stage, the largest error
against the oracle evaluated in float64 (relative to the stage's largest magnitude), beside the same figure for the
unmodified reference's fp32 CPU outputs (the fixture).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
from conftest import load_golden, make_case                     # noqa: E402
from oracle import regtr_oracle as O                             # noqa: E402
from regtr_b200 import kpconv as K                               # noqa: E402
from regtr_b200.regtr import RegTR                               # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else 'var_modelnet_learnedpe_attndec_b2'
cfg, sd, src, tgt = make_case(case)
fx = load_golden(case)
o64 = O.forward(sd, cfg, src, tgt, dtype=torch.float64)
KEYS = ['src_feat_un', 'src_feat', 'src_kp_warped', 'src_overlap']


def run(label, attention='fp32', **env):
    for k, v in env.items():
        os.environ[k] = v
    c = cfg.copy() if hasattr(cfg, 'copy') else cfg
    c.attention_impl = attention
    model = RegTR(c).to('cuda:0').eval()
    model.load_state_dict(sd, strict=True)
    out = model({'src_xyz': [torch.from_numpy(a).cuda() for a in src], 'tgt_xyz': [torch.from_numpy(a).cuda() for a in tgt]})
    torch.cuda.synchronize()
    for k in env:
        del os.environ[k]
    row = []
    for key in KEYS:
        e = 0.0
        for b in range(len(src)):
            w = o64[key][b].numpy()
            g = out[key][b].detach().cpu().numpy()
            e = max(e, np.abs(g - w).max() / max(np.abs(w).max(), 1e-30))
        row.append(e)
    pose = out['pose'].detach().cpu().numpy()
    print(f'{label:34s} ' + ' '.join(f'{k} {e:.2e}' for k, e in zip(KEYS, row)) +
          f' | pose vs f64 {np.abs(pose - o64["pose"].numpy()).max():.3e} vs ref {np.abs(pose - fx["pose"]).max():.3e}')


print(case, 'reference fp32 CPU pose vs f64: %.3e' % np.abs(fx['pose'] - o64['pose'].numpy()).max())
run('default')
K.EPILOGUE_STATS = False
run('stand-alone IN statistics')
K.EPILOGUE_STATS = True
run('agg=ffma', REGTR_AGG_IMPL='ffma')
run('agg=mma', REGTR_AGG_IMPL='mma')
run('mha=ffma', REGTR_MHA_IMPL='ffma')
run('gemm=ffma', REGTR_GEMM_IMPL='ffma')
run('gemm=ffma + agg=ffma', REGTR_GEMM_IMPL='ffma', REGTR_AGG_IMPL='ffma')
run('attention tf32_tc', attention='tf32_tc')
K.EPILOGUE_STATS = False
run('stats off + all ffma', REGTR_AGG_IMPL='ffma', REGTR_MHA_IMPL='ffma', REGTR_GEMM_IMPL='ffma')
