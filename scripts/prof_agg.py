"""Minimal driver for ncu: the aggregation kernel on levels 0..3 of one synthetic 3DMatch pair."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regtr_b200 import ops, config, synthetic
from regtr_b200.kpconv import PreprocessorGPU

dev = torch.device('cuda:0')
cfg = config.regtr_3dmatch()
pair = synthetic.make_3dmatch_pair(2000)
meta = PreprocessorGPU(cfg)([torch.from_numpy(pair['src_xyz']).to(dev), torch.from_numpy(pair['tgt_xyz']).to(dev)])
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for lvl, cin in ((0, 32), (1, 64), (2, 128), (3, 256)):
    pts = meta['points'][lvl]; idx = meta['neighbors'][lvl].to(torch.int32)
    r = 0.0625 * 2 ** lvl
    x = torch.relu(torch.randn(pts.shape[0], cin, device=dev))
    kp = torch.randn(15, 3, device=dev) * r * 0.5
    flags = torch.ones(pts.shape[0], dtype=torch.uint8, device=dev)
    for _ in range(2):
        flush.zero_()
        ops.kpconv_aggregate(pts, pts, idx, x, kp, 0.8 * r, row_flags=flags)
torch.cuda.synchronize()
