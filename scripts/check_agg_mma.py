"""A/B of the two KPConv aggregation kernels (tensor-core mma.sync vs packed-FFMA) on the real pyramid
of a synthetic 3DMatch pair: max abs difference against a float64 torch restatement and per-launch time."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regtr_b200 import ops, config, synthetic
from regtr_b200.kpconv import PreprocessorGPU

dev = torch.device('cuda:0')
cfg = config.regtr_3dmatch()
pair = synthetic.make_3dmatch_pair(2000)
pre = PreprocessorGPU(cfg)
meta = pre([torch.from_numpy(pair['src_xyz']).to(dev), torch.from_numpy(pair['tgt_xyz']).to(dev)])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def ref64(q, s, idx, x, kp, extent):
    s1 = torch.cat([s, torch.full((1, 3), 1e6, device=dev)]).double()
    x1 = torch.cat([x, torch.zeros(1, x.shape[1], device=dev)]).double()
    out = torch.empty(q.shape[0], 15, x.shape[1], dtype=torch.float64, device=dev)
    for a in range(0, q.shape[0], 4096):
        i = idx[a:a + 4096].long()
        nb = s1[i] - q[a:a + 4096, None].double()
        d = (nb[:, :, None, :] - kp.double()[None, None]).norm(dim=-1)
        w = (1 - d / extent).clamp(min=0)
        nx = x1[i]
        cnt = (nx.sum(-1) > 0).sum(-1).clamp(min=1)
        out[a:a + 4096] = torch.einsum('nkp,nkc->npc', w, nx) / cnt[:, None, None]
    return out

torch.manual_seed(0)
for lvl, cin in ((0, 1), (0, 32), (1, 64), (2, 128), (3, 256)):
    pts = meta['points'][lvl]; idx = meta['neighbors'][lvl].to(torch.int32)
    r = 0.0625 * 2 ** lvl
    x = torch.relu(torch.randn(pts.shape[0], cin, device=dev)) * (torch.rand(pts.shape[0], 1, device=dev) > 0.05)
    kp = torch.randn(15, 3, device=dev) * r * 0.5
    kp[0] = 0
    ref = ref64(pts, pts, idx, x, kp, 0.8 * r)
    for impl, mw in (('ffma', 0), ('mma', 0), ('mma', 4096), ('mma', 8192), ('mma', 1 << 30)):
        os.environ['REGTR_AGG_IMPL'] = impl
        os.environ['REGTR_AGG_MIN_WARPS'] = str(mw)
        wf = ops.kpconv_aggregate(pts, pts, idx, x, kp, 0.8 * r).view(-1, 15, cin)
        err = (wf.double() - ref).abs().max().item()
        best = 1e9
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.kpconv_aggregate(pts, pts, idx, x, kp, 0.8 * r); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3)
        print(f'level {lvl} Nq={pts.shape[0]} Cin={cin} {impl:5s} min_warps={mw:<10d} max|err|={err:.3e} (ref max {ref.abs().max().item():.3f})  {best:.1f} us', flush=True)
