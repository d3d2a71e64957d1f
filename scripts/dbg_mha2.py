import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regtr_b200 import ops
from regtr_b200.transformer import AttentionPlan
dev = 'cuda:0'
E, H = 256, 8
lens = [int(v) for v in os.environ.get("LENS", "64,64").split(",")]
N = sum(lens)
x = torch.randn(N, E, device=dev); W = torch.randn(3 * E, E, device=dev) / 16; b = torch.zeros(3 * E, device=dev)
plan = AttentionPlan(lens, dev)
torch.cuda.synchronize()
print('calling', flush=True)
got = ops.mha_bf16_tc(x, W, b, plan.q_start, plan.q_len, plan.q_start, plan.q_len, plan.max_len, H)
torch.cuda.synchronize()
print('done', float(got.abs().max()))
