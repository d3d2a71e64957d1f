"""Diagnose the fused Cin=1 KPConv kernel against float64 (where is the error, fused vs aggregate-only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from regtr_b200 import ops
DEV = 'cuda:0'
G = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(DEV) if dt is None else torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dt)
cin, cout = 1, 64
rng = np.random.default_rng(cin)
Nq, Ns, K = 301, 457, 40
q = rng.normal(size=(Nq, 3)).astype(np.float32) * 0.05
s = rng.normal(size=(Ns, 3)).astype(np.float32) * 0.05
idx = rng.integers(0, Ns + 1, size=(Nq, K))
idx[:, 30:] = np.where(rng.random((Nq, 10)) < 0.7, Ns, idx[:, 30:])
idx[7] = Ns
x = rng.normal(size=(Ns, cin)).astype(np.float32) + 1.0
W = (rng.normal(size=(15, cin, cout)) / np.sqrt(15 * cin)).astype(np.float32)
kp = (rng.normal(size=(15, 3)) * 0.03).astype(np.float32)
s64 = np.concatenate([s.astype(np.float64), np.full((1, 3), 1e6)])
x64 = np.concatenate([x.astype(np.float64), np.zeros((1, cin))])
nb = s64[idx] - q[:, None, :].astype(np.float64)
d = np.linalg.norm(nb[:, :, None, :] - kp[None, None].astype(np.float64), axis=-1)
h = np.clip(1 - d / np.float64(np.float32(0.05)), 0, None)
nx = x64[idx]
wf = np.einsum('nkp,nkc->npc', h, nx)
cnt = np.maximum((nx.sum(-1) > 0).sum(1), 1)
wfn = wf[:, :, 0] / cnt[:, None]
out = wfn @ W[:, 0, :].astype(np.float64)
got = ops.kpconv(G(q), G(s), G(idx, torch.int32), G(x), G(W), G(kp), 0.05).cpu().numpy()
err = np.abs(got - out)
print('fused: max err', err.max(), 'at', np.unravel_index(err.argmax(), err.shape), 'scale', np.abs(out).max(),
      'rows with err > 2e-6:', int((err.max(1) > 2e-6).sum()))
agg = ops.kpconv_aggregate(G(q), G(s), G(idx, torch.int32), G(x), G(kp), 0.05).cpu().numpy()
e2 = np.abs(agg - wfn)
print('aggregate-only: max err', e2.max(), 'at', np.unravel_index(e2.argmax(), e2.shape), 'scale', np.abs(wfn).max())
bad = np.argsort(-err.max(1))[:5]
for r in bad:
    print('row', r, 'err', err[r].max(), 'n_valid', int((idx[r] < Ns).sum()), 'cnt', cnt[r], 'agg err', e2[r].max(),
          'min h>0', h[r][h[r] > 0].min() if (h[r] > 0).any() else None)
