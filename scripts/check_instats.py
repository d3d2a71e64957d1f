"""GEMM with the InstanceNorm statistics epilogue vs float64 (diagnostics: poisoned = row-count mismatch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from regtr_b200 import ops
DEV = 'cuda:0'
G = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
ok = True
for lens, N, K in [([700, 1, 0, 333, 90], 64, 96), ([4000, 4100], 128, 64), ([300, 260], 256, 3840), ([5000, 4000, 3000, 100], 32, 480),
                   ([128, 128], 128, 32), ([1000], 32, 64), ([20000, 18000], 128, 64), ([300, 227], 1024, 512)]:
    rng = np.random.default_rng(N + K)
    M = sum(lens); cap = M + 200
    a = np.zeros((cap, K), dtype=np.float32); a[:M] = rng.normal(size=(M, K)) * 1.3 + 0.4; a[M:] = 1e3
    w = (rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)
    offs = ops.make_offsets(lens, DEV); m_dev = offs[len(lens):len(lens) + 1]
    hi, lo = ops.split_weight(G(w))
    res = [ops.gemm_instats(G(a), hi, lo, offs, len(lens), m_dev=m_dev) for _ in range(3)]
    torch.cuda.synchronize()
    c64 = a[:M].astype(np.float64) @ w.astype(np.float64).T
    st = res[0][1].cpu().numpy()
    det = all(torch.equal(res[0][1], r[1]) and torch.equal(res[0][0][:M], r[0][:M]) for r in res[1:])
    starts = np.concatenate([[0], np.cumsum(lens)])
    em = er = 0.0
    for c, n in enumerate(lens):
        if n == 0: continue
        blk = c64[starts[c]:starts[c + 1]]
        em = max(em, np.nanmax(np.abs(st[c, :, 0] - blk.mean(0))))
        er = max(er, np.nanmax(np.abs(st[c, :, 1] * np.sqrt(blk.var(0) + 1e-5) - 1)))
    eo = np.abs(res[0][0].cpu().numpy()[:M] - c64).max() / np.abs(c64).max()
    good = det and not np.isnan(st).any() and em < 5e-6 * max(1, np.abs(c64).max()) and er < 3e-5 and eo < 2e-5
    ok &= good
    print(f'lens {lens} N {N} K {K}: deterministic {det} poisoned {int(np.isnan(st).sum())} mean err {em:.2e} rstd rel err {er:.2e} out rel err {eo:.2e} {"OK" if good else "FAIL"}')
print('ALL OK' if ok else 'SOME FAILED')
