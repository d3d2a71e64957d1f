"""Attention-only loop for ncu: B pairs of (410, 339) tokens, self + cross, both implementations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regtr_b200 import ops
from regtr_b200.transformer import AttentionPlan
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = 'cuda:0'; E, H = 256, 8
L = [410] * B + [339] * B
x = torch.randn(sum(L), E, device=dev); W = torch.randn(3 * E, E, device=dev) / 16; b = torch.zeros(3 * E, device=dev)
plan = AttentionPlan(L, dev)
qkv = ops.linear(x, W, b)
for _ in range(3):
    ops.mha_varlen(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], plan.q_start, plan.q_len, plan.xk_start, plan.xk_len, plan.max_len, H)
    ops.mha_bf16_tc(x, W, b, plan.q_start, plan.q_len, plan.xk_start, plan.xk_len, plan.max_len, H)
torch.cuda.synchronize(); print('ok')
