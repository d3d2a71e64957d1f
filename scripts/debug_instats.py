"""Debug of the statistics epilogue: constant rows (C[r, j] = colsum(w[j]) for every row) make every
contribution identical, so mean / expected = (#rows counted) / n exposes lost or extra contributions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from regtr_b200 import ops
DEV = 'cuda:0'
G = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
for lens, N, K, use_mdev, pad in [([128, 128], 128, 32, True, 200), ([128, 128], 128, 32, False, 0), ([128, 128], 128, 32, True, 0),
                                  ([256], 64, 32, True, 200), ([256], 32, 32, True, 200), ([256], 32, 32, False, 0),
                                  ([1000], 32, 64, True, 200), ([1000], 32, 64, False, 0), ([1000], 64, 64, True, 200)]:
    M = sum(lens); cap = M + pad
    a = np.ones((cap, K), dtype=np.float32); a[M:] = 1e3
    rng = np.random.default_rng(1)
    w = (rng.integers(-3, 4, size=(N, K))).astype(np.float32)          # small integers: every product exact
    offs = ops.make_offsets(lens, DEV); m_dev = offs[len(lens):len(lens) + 1] if use_mdev else None
    hi, lo = ops.split_weight(G(w))
    out, stats = ops.gemm_instats(G(a), hi, lo, offs, len(lens), m_dev=m_dev)
    torch.cuda.synchronize()
    col = w.sum(1)                                                      # C[r, j] for every valid row
    st = stats.cpu().numpy()
    key = [k for k in ops._ws_cache if k[2] == 'instnorm_part'][0]
    accbuf = ops._ws_cache[key]
    nz = int((accbuf != 0).sum())
    okc = np.abs(out.cpu().numpy()[:M] - col[None]).max()
    ratio = st[:, :, 0] / np.where(col == 0, 1, col)[None]
    print(f'lens {lens} N {N} K {K} m_dev {use_mdev} pad {pad}: out err {okc:.1e}; acc nonzero bytes after {nz}; '
          f'mean/expected min {np.nanmin(ratio[:, col != 0]):.4f} max {np.nanmax(ratio[:, col != 0]):.4f}; '
          f'rstd min {st[:, :, 1].min():.3f} max {st[:, :, 1].max():.3f} (expect 316.228)')
    bad = np.argwhere(np.abs(ratio - 1) > 1e-3)
    if len(bad):
        print('   bad (cloud, col) first 12:', bad[:12].tolist(), ' n bad', len(bad), 'of', ratio.size)
