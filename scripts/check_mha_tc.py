import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regtr_b200 import ops
from regtr_b200.transformer import AttentionPlan
torch.manual_seed(0)
dev = 'cuda:0'
E, H = 256, 8
ok = True
for lens in ([410, 339], [64, 64], [130, 7, 300, 129], [650, 600]):
    N = sum(lens)
    x = torch.randn(N, E, device=dev)
    W = torch.randn(3 * E, E, device=dev) / E ** 0.5
    b = torch.randn(3 * E, device=dev) * 0.1
    plan = AttentionPlan(lens, dev)
    for cross in (False, True):
        ks, kl = (plan.xk_start, plan.xk_len) if cross else (plan.q_start, plan.q_len)
        qkv = ops.linear(x, W, b)
        want = ops.mha_varlen(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], plan.q_start, plan.q_len, ks, kl, plan.max_len, H)
        got = ops.mha_bf16_tc(x, W, b, plan.q_start, plan.q_len, ks, kl, plan.max_len, H)
        torch.cuda.synchronize()
        err = float((got - want).abs().max()); scale = float(want.abs().max())
        good = err <= 3e-2 * scale and bool(torch.isfinite(got).all())
        ok &= good
        print(f'lens {lens} cross {cross}: err {err:.3e} scale {scale:.3f} rel {err/scale:.2e} {"OK" if good else "FAIL"}')
# timing
lens = [410, 339] * 1
for B in (1, 8):
    L = ([410] * B) + ([339] * B)
    N = sum(L)
    x = torch.randn(N, E, device=dev); W = torch.randn(3 * E, E, device=dev) / 16; b = torch.zeros(3 * E, device=dev)
    plan = AttentionPlan(L, dev)
    qkv = ops.linear(x, W, b)
    for _ in range(3):
        ops.mha_varlen(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], plan.q_start, plan.q_len, plan.xk_start, plan.xk_len, plan.max_len, H)
        ops.mha_bf16_tc(x, W, b, plan.q_start, plan.q_len, plan.xk_start, plan.xk_len, plan.max_len, H)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(20):
        ops.mha_varlen(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], plan.q_start, plan.q_len, plan.xk_start, plan.xk_len, plan.max_len, H)
    e[1].record()
    for _ in range(20):
        ops.mha_bf16_tc(x, W, b, plan.q_start, plan.q_len, plan.xk_start, plan.xk_len, plan.max_len, H)
    e[2].record(); torch.cuda.synchronize()
    print(f'B={B}: fp32 core {e[0].elapsed_time(e[1])/20*1e3:.1f} us ; bf16_tc (in-proj GEMM + core) {e[1].elapsed_time(e[2])/20*1e3:.1f} us')
print('ALL OK' if ok else 'SOME FAILED')
