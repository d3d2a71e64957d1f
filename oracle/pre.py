"""ORACLE (test infrastructure only) -- pyramid pre-processing on the CPU.

`ball_query` / `grid_subsample` bind the plain-C restatement in
oracle/c/preprocess_oracle.c; the `_np` variants are the same rules written
with numpy for tiny inputs (they cross-check the C build in tests/).
`preprocess` restates PreprocessorGPU.forward
(/root/reference/src/models/backbone_kpconv/kpconv.py:426-537) on top of them.

`RefCpp` binds oracle/_ref/libregtr_ref_cpp.so -- the reference's own C++ core
(nanoflann radius search, hash-map grid subsampling) -- and `preprocess_refcpp`
restates the reference's CPU `Preprocessor.forward` (kpconv.py:291-414) on it;
that is the pre-processing leg of the CPU baseline, not the index-parity oracle.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build(ref: bool = True) -> None:
    """Compile the C restatement (and, when /root/reference exists, oracle/_ref)."""
    subprocess.run(['make', '-C', _HERE, 'all'], check=True, capture_output=True)
    if ref and os.path.isdir('/root/reference/src'):
        subprocess.run(['make', '-C', _HERE, 'ref'], check=True, capture_output=True)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'liboracle_pre.so')
        if not os.path.exists(path):
            build(ref=False)
        lib = ctypes.CDLL(path)
        f32p, i64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        lib.oracle_ball_query.argtypes = [f32p, i64p, f32p, i64p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_float, i64p]
        lib.oracle_ball_query.restype = ctypes.c_int
        lib.oracle_grid_subsample.argtypes = [f32p, i64p, ctypes.c_int, ctypes.c_float, f32p, i64p]
        lib.oracle_grid_subsample.restype = ctypes.c_int64
        _LIB = lib
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def ball_query(q, q_lens, s, s_lens, K: int, radius: float) -> np.ndarray:
    """(Nq,K) int64 packed support indices, first K in index order, pad = Ns_total."""
    q, s, ql, sl = _f32(q), _f32(s), _i64(q_lens), _i64(s_lens)
    out = np.empty((q.shape[0], K), dtype=np.int64)
    rc = _lib().oracle_ball_query(_p(q, ctypes.c_float), _p(ql, ctypes.c_int64),
                                  _p(s, ctypes.c_float), _p(sl, ctypes.c_int64),
                                  len(ql), K, np.float32(radius), _p(out, ctypes.c_int64))
    assert rc == 0
    return out


def grid_subsample(xyz, lens, dl: float):
    """-> (sub_xyz (M,3) f32, sub_lens (C,) i64) in ascending (cloud, vx, vy, vz) order."""
    xyz, lens = _f32(xyz), _i64(lens)
    out = np.empty_like(xyz)
    out_lens = np.zeros_like(lens)
    m = _lib().oracle_grid_subsample(_p(xyz, ctypes.c_float), _p(lens, ctypes.c_int64), len(lens),
                                     np.float32(dl), _p(out, ctypes.c_float),
                                     _p(out_lens, ctypes.c_int64))
    assert m >= 0
    return out[:m].copy(), out_lens


# ------------------------------------------------------------------ numpy twins

def ball_query_np(q, q_lens, s, s_lens, K, radius):
    q, s = _f32(q), _f32(s)
    r = np.float32(radius)
    r2 = r * r
    out = np.full((q.shape[0], K), s.shape[0], dtype=np.int64)
    q0 = s0 = 0
    for nq, ns in zip(map(int, q_lens), map(int, s_lens)):
        d = q[q0:q0 + nq, None, :] - s[None, s0:s0 + ns, :]
        sq = d * d
        d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
        hit = d2 < r2
        for i in range(nq):
            js = np.nonzero(hit[i])[0][:K]
            out[q0 + i, :len(js)] = js + s0
        q0 += nq
        s0 += ns
    return out


def grid_subsample_np(xyz, lens, dl):
    xyz = _f32(xyz)
    dl32 = np.float32(dl)
    vox = np.floor(xyz / dl32).astype(np.int64)
    cloud = np.repeat(np.arange(len(lens)), np.asarray(lens, dtype=np.int64))
    order = np.lexsort((np.arange(len(xyz)), vox[:, 2], vox[:, 1], vox[:, 0], cloud))
    keys = np.concatenate([cloud[order, None], vox[order]], axis=1)
    new = np.ones(len(order), dtype=bool)
    new[1:] = np.any(keys[1:] != keys[:-1], axis=1)
    starts = np.nonzero(new)[0]
    ends = np.append(starts[1:], len(order))
    out = np.empty((len(starts), 3), dtype=np.float32)
    out_lens = np.zeros(len(lens), dtype=np.int64)
    for m, (a, b) in enumerate(zip(starts, ends)):
        acc = np.zeros(3, dtype=np.float32)
        for i in order[a:b]:
            acc = acc + xyz[i]
        out[m] = acc / np.float32(b - a)
        out_lens[keys[a, 0]] += 1
    return out, out_lens


# --------------------------------------------------------------------- pyramid

def preprocess(cfg, pts_list, with_upsamples: bool = True):
    """Oracle twin of PreprocessorGPU.forward (kpconv.py:426-537); numpy in/out.

    Returns dict with lists `points`, `neighbors`, `pools`, `upsamples`,
    `stack_lengths` (np arrays; empty pools/upsamples on the last level have the
    reference's (0,1) / (0,3) / (0,) shapes).
    """
    from regtr_b200.config import pyramid_plan
    levels, _, _ = pyramid_plan(cfg)
    pts = _f32(np.concatenate([np.asarray(p) for p in pts_list], axis=0))
    lens = np.array([len(p) for p in pts_list], dtype=np.int64)
    out = dict(points=[], neighbors=[], pools=[], upsamples=[], stack_lengths=[])
    for lvl in levels:
        r, K = lvl['radius'], lvl['K']
        conv_i = ball_query(pts, lens, pts, lens, K, r) if lvl['has_conv'] \
            else np.zeros((0, 1), np.int64)
        if lvl['strided']:
            pool_p, pool_b = grid_subsample(pts, lens, lvl['dl'])
            pool_i = ball_query(pool_p, pool_b, pts, lens, K, r)
            up_i = ball_query(pts, lens, pool_p, pool_b, K, 2 * r) if with_upsamples \
                else np.zeros((0, 1), np.int64)
        else:
            pool_i = np.zeros((0, 1), np.int64)
            pool_p = np.zeros((0, 3), np.float32)
            pool_b = np.zeros((0,), np.int64)
            up_i = np.zeros((0, 1), np.int64)
        out['points'].append(pts)
        out['neighbors'].append(conv_i)
        out['pools'].append(pool_i)
        out['upsamples'].append(up_i)
        out['stack_lengths'].append(lens)
        pts, lens = pool_p, pool_b
    return out


# ------------------------------------------------------- reference C++ (ORACLE-C)

class RefCpp:
    """ctypes binding of oracle/_ref/libregtr_ref_cpp.so (reference's C++ core)."""

    def __init__(self):
        path = os.path.join(_HERE, '_ref', 'libregtr_ref_cpp.so')
        if not os.path.exists(path):
            raise FileNotFoundError(f'{path} missing: run `make -C oracle ref` where /root/reference exists')
        lib = ctypes.CDLL(path)
        f32p, i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
        lib.regtr_ref_batch_neighbors.argtypes = [f32p, ctypes.c_int64, f32p, ctypes.c_int64, i32p,
                                                  i32p, ctypes.c_int, ctypes.c_float, i32p, ctypes.c_int]
        lib.regtr_ref_batch_neighbors.restype = ctypes.c_int
        lib.regtr_ref_batch_subsample.argtypes = [f32p, ctypes.c_int64, i32p, ctypes.c_int,
                                                  ctypes.c_float, f32p, i32p]
        lib.regtr_ref_batch_subsample.restype = ctypes.c_int64
        self.lib = lib

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(_HERE, '_ref', 'libregtr_ref_cpp.so'))

    def batch_neighbors(self, q, s, q_lens, s_lens, radius, K):
        """K nearest within radius, sorted by distance (kpconv.py:243-258 semantics)."""
        q, s = _f32(q), _f32(s)
        ql = np.ascontiguousarray(q_lens, dtype=np.int32)
        sl = np.ascontiguousarray(s_lens, dtype=np.int32)
        out = np.empty((q.shape[0], K), dtype=np.int32)
        self.lib.regtr_ref_batch_neighbors(_p(q, ctypes.c_float), q.shape[0], _p(s, ctypes.c_float),
                                           s.shape[0], _p(ql, ctypes.c_int32), _p(sl, ctypes.c_int32),
                                           len(ql), np.float32(radius), _p(out, ctypes.c_int32), K)
        return out.astype(np.int64)

    def batch_subsample(self, xyz, lens, dl):
        xyz = _f32(xyz)
        ln = np.ascontiguousarray(lens, dtype=np.int32)
        out = np.empty_like(xyz)
        out_lens = np.zeros_like(ln)
        m = self.lib.regtr_ref_batch_subsample(_p(xyz, ctypes.c_float), xyz.shape[0],
                                               _p(ln, ctypes.c_int32), len(ln), np.float32(dl),
                                               _p(out, ctypes.c_float), _p(out_lens, ctypes.c_int32))
        return out[:m].copy(), out_lens.astype(np.int64)


def preprocess_refcpp(cfg, pts_list, ref: RefCpp | None = None):
    """Twin of the reference's CPU Preprocessor.forward (kpconv.py:291-414) on its C++ core."""
    from regtr_b200.config import pyramid_plan
    ref = ref or RefCpp()
    levels, _, _ = pyramid_plan(cfg)
    pts = _f32(np.concatenate([np.asarray(p) for p in pts_list], axis=0))
    lens = np.array([len(p) for p in pts_list], dtype=np.int64)
    out = dict(points=[], neighbors=[], pools=[], upsamples=[], stack_lengths=[])
    for lvl in levels:
        r, K = lvl['radius'], lvl['K']
        conv_i = ref.batch_neighbors(pts, pts, lens, lens, r, K) if lvl['has_conv'] \
            else np.zeros((0, 1), np.int64)
        if lvl['strided']:
            pool_p, pool_b = ref.batch_subsample(pts, lens, lvl['dl'])
            pool_i = ref.batch_neighbors(pool_p, pts, pool_b, lens, r, K)
            up_i = ref.batch_neighbors(pts, pool_p, lens, pool_b, 2 * r, K)
        else:
            pool_i = np.zeros((0, 1), np.int64)
            pool_p = np.zeros((0, 3), np.float32)
            pool_b = np.zeros((0,), np.int64)
            up_i = np.zeros((0, 1), np.int64)
        out['points'].append(pts)
        out['neighbors'].append(conv_i)
        out['pools'].append(pool_i)
        out['upsamples'].append(up_i)
        out['stack_lengths'].append(lens)
        pts, lens = pool_p, pool_b
    return out
