"""A/B of the two fp32-accurate attention cores (mma.sync 3xTF32 vs CUDA-core FFMA) against float64."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regtr_b200 import ops
from regtr_b200.transformer import AttentionPlan

dev = 'cuda:0'; E, H = 256, 8
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for L in ([410, 339], [410] * 8 + [339] * 8, [1, 7], [64, 65], [130, 3]):
    B = len(L) // 2
    x = torch.randn(sum(L), 3 * E, device=dev) * 1.5
    q, k, v = x[:, :E], x[:, E:2 * E], x[:, 2 * E:]
    plan = AttentionPlan(L, dev)
    for name, ks, kn in (('self', plan.q_start, plan.q_len), ('cross', plan.xk_start, plan.xk_len)):
        # float64 reference
        ref = torch.zeros(sum(L), E, dtype=torch.float64, device=dev)
        qs, qn, kss, knn = plan.q_start.tolist(), plan.q_len.tolist(), ks.tolist(), kn.tolist()
        for p in range(len(L)):
            qq = q[qs[p]:qs[p] + qn[p]].double().view(-1, H, 32).transpose(0, 1)
            kk = k[kss[p]:kss[p] + knn[p]].double().view(-1, H, 32).transpose(0, 1)
            vv = v[kss[p]:kss[p] + knn[p]].double().view(-1, H, 32).transpose(0, 1)
            a = torch.softmax(qq @ kk.transpose(1, 2) / 32 ** 0.5, -1) @ vv
            ref[qs[p]:qs[p] + qn[p]] = a.transpose(0, 1).reshape(-1, E)
        for impl in ('ffma', 'mma'):
            os.environ['REGTR_MHA_IMPL'] = impl
            o = ops.mha_varlen(q, k, v, plan.q_start, plan.q_len, ks, kn, plan.max_len, H)
            err = (o.double() - ref).abs().max().item()
            best = 1e9
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ops.mha_varlen(q, k, v, plan.q_start, plan.q_len, ks, kn, plan.max_len, H); e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e3)
            print(f'L={L[:2]}x{B} {name:5s} {impl:4s} max|err|={err:.3e}  {best:.1f} us', flush=True)
