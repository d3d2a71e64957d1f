"""Torch-tensor front end of the C ABI (one function per entry point of include/regtr_b200.h).

PyTorch is used for device memory and streams only; every operation below runs a
hand-written sm_100a kernel from libregtr_b200.so on the current CUDA stream.
Tensors must live on a CUDA device; there is no CPU fallback.
"""
from __future__ import annotations

import contextlib
import itertools
import math

import torch

from . import lib as _lib

_ws_cache = {}

# Number of hand-written kernels (libregtr_b200.so, excluding CUB / cuBLAS) launched so far.
LAUNCHES = 0
# Optional profiler hooks (bench.py only): when KPCONV_TRACE is a list, ops.kpconv appends (cuda start event,
# end event, info dict) around every KPConv call; when TRACE is a list, the GEMM / attention-core / gather
# front ends append (kind, info dict, re-launch closure) so that the bench can re-time every launch of one
# forward on its own (L2 flushed) and build per-kernel-family rooflines from the real shapes.
KPCONV_TRACE = None
TRACE = None


def _count(n):
    global LAUNCHES
    LAUNCHES += n


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(t, dtype, name, dims=None):
    if not t.is_cuda:
        raise _lib.RegtrLibError(f'{name}: expected a CUDA tensor (the product path has no CPU fallback)')
    if t.dtype != dtype:
        raise TypeError(f'{name}: expected {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError(f'{name}: must be contiguous')
    if dims is not None and t.dim() != dims:
        raise ValueError(f'{name}: expected {dims}-d tensor, got shape {tuple(t.shape)}')
    return t


# Scratch buffers.  Eager calls share one set per (device, CUDA stream): stream order makes the reuse
# safe and two streams never alias each other's scratch (nor the self-resetting InstanceNorm counters).
# A CUDA-graph capture runs under its own NAMESPACE token (`scratch_namespace`): the raw pointers baked into
# that graph then belong to that graph alone; a buffer that has to grow inside a namespace keeps its
# predecessor alive (`_ws_retired`) because an already captured graph may still write to it, and everything
# is released together with `release_namespace` when the graph is dropped.
WS_NAMESPACE = None
_ws_retired = {}
_ns_counter = itertools.count(1)


def new_namespace():
    return ('graph', next(_ns_counter))


@contextlib.contextmanager
def scratch_namespace(ns):
    global WS_NAMESPACE
    prev, WS_NAMESPACE = WS_NAMESPACE, ns
    try:
        yield ns
    finally:
        WS_NAMESPACE = prev


def release_namespace(ns):
    """Drop every scratch buffer of a namespace (call when its captured graph is destroyed)."""
    for key in [k for k in _ws_cache if k[1] == ns]:
        del _ws_cache[key]
    _ws_retired.pop(ns, None)


def workspace(nbytes: int, device, slot: str = 'default', zero: bool = False) -> torch.Tensor:
    """Per-(device, namespace | stream, slot) grow-only scratch buffer (stream-ordered reuse).
    zero=True: allocated zero-filled (state that an op keeps zero between its own calls)."""
    dev = device.index if device.index is not None else torch.cuda.current_device()
    ns = WS_NAMESPACE if WS_NAMESPACE is not None else ('stream', torch.cuda.current_stream(device).cuda_stream)
    key = (dev, ns, slot)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None and WS_NAMESPACE is not None:
            _ws_retired.setdefault(ns, []).append(buf)      # a captured graph may hold this pointer
        n = max(int(nbytes * 1.25), 4096 if zero else 1 << 20)
        buf = (torch.zeros if zero else torch.empty)(n, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def regtr_align_up(n, a=256):
    return (n + a - 1) // a * a


def make_offsets(lengths, device) -> torch.Tensor:
    """int32 prefix offsets (n_clouds+1) on `device` from a host list / tensor of lengths."""
    if torch.is_tensor(lengths):
        lengths = lengths.tolist()
    offs = [0]
    for v in lengths:
        offs.append(offs[-1] + int(v))
    return torch.tensor(offs, dtype=torch.int32, device=device)


def new_status(device) -> torch.Tensor:
    return torch.zeros(1, dtype=torch.int32, device=device)


# ---------------------------------------------------------------- pre-processing

def grid_subsample(xyz, offs, n_clouds: int, dl: float, status, out_cap=None, out_offs=None, dense: bool = True):
    """-> (out_xyz (out_cap,3) capacity buffer, out_offs (n_clouds+1) int32).  No host sync.
    out_cap defaults to the input capacity (always sufficient); a smaller value is memory-safe but
    raises REGTR_STATUS_CAPACITY in `status` when the sub-sampled level does not fit.
    dense=True: counting sort over a dense voxel grid (hand-written kernels; raises REGTR_STATUS_GRID in
    `status` when a cloud's bounding box exceeds the cell budget); dense=False: sort-based variant (any extent)."""
    L = _lib.load()
    _chk(xyz, torch.float32, 'xyz', 2); _chk(offs, torch.int32, 'offs', 1)
    n_cap = xyz.shape[0]
    out_cap = n_cap if out_cap is None else int(out_cap)
    out_xyz = torch.empty((out_cap, 3), dtype=torch.float32, device=xyz.device)
    out_offs = torch.empty(n_clouds + 1, dtype=torch.int32, device=xyz.device) if out_offs is None else out_offs
    if dense:
        ws = workspace(L.regtr_grid_subsample_ws_bytes(n_cap, n_clouds), xyz.device)
        state = workspace(L.regtr_grid_subsample_state_bytes(n_cap), xyz.device, 'subsample_state', zero=True)
        _lib.check(L.regtr_grid_subsample(_p(xyz), _p(offs), n_clouds, n_cap, float(dl), _p(out_xyz), out_cap,
                                          _p(out_offs), _p(status), _p(ws), ws.numel(), _p(state), state.numel(),
                                          _stream()), 'regtr_grid_subsample')
        _count(7)
    else:
        ws = workspace(L.regtr_grid_subsample_sorted_ws_bytes(n_cap), xyz.device)
        _lib.check(L.regtr_grid_subsample_sorted(_p(xyz), _p(offs), n_clouds, n_cap, float(dl), _p(out_xyz), out_cap,
                                                 _p(out_offs), _p(status), _p(ws), ws.numel(), _stream()),
                   'regtr_grid_subsample_sorted')
        _count(4)
    return out_xyz, out_offs


class CellGrid:
    """Opaque cell list over a stacked point set (+ the cell-sorted point order)."""

    def __init__(self, xyz, offs, n_clouds: int, cell: float, status):
        L = _lib.load()
        _chk(xyz, torch.float32, 'xyz', 2); _chk(offs, torch.int32, 'offs', 1)
        self.n_cap = xyz.shape[0]
        self.n_clouds = n_clouds
        self.cell = float(cell)
        self.buf = torch.empty(L.regtr_cellgrid_bytes(self.n_cap), dtype=torch.uint8, device=xyz.device)
        self.order = torch.empty(max(self.n_cap, 1), dtype=torch.int32, device=xyz.device)
        ws = workspace(L.regtr_cellgrid_ws_bytes(self.n_cap), xyz.device)
        state = workspace(L.regtr_cellgrid_state_bytes(self.n_cap), xyz.device, 'scan_state', zero=True)
        _lib.check(L.regtr_cellgrid_build(_p(xyz), _p(offs), n_clouds, self.n_cap, self.cell, _p(self.buf),
                                          _p(self.order), _p(status), _p(ws), ws.numel(), _p(state), state.numel(),
                                          _stream()), 'regtr_cellgrid_build')
        _count(4)


def ball_query(q, q_offs, s, s_offs, grid: CellGrid, K: int, radius: float, q_order=None,
               want32=True, want64=True):
    """First-K-in-index-order radius search.  -> (idx32 or None, idx64 or None), shape (nq_cap,K)."""
    L = _lib.load()
    _chk(q, torch.float32, 'q', 2); _chk(s, torch.float32, 's', 2)
    nq_cap = q.shape[0]
    if grid.n_cap != s.shape[0]:
        raise ValueError('grid was built over a different support capacity')
    if grid.cell < float(radius):
        raise ValueError('grid cell must be >= radius')
    i32 = torch.empty((nq_cap, K), dtype=torch.int32, device=q.device) if want32 else None
    i64 = torch.empty((nq_cap, K), dtype=torch.int64, device=q.device) if want64 else None
    _lib.check(L.regtr_ball_query(_p(q), _p(q_offs), _p(q_order), _p(s), _p(s_offs), _p(grid.buf), grid.n_clouds,
                                  nq_cap, s.shape[0], int(K), float(radius), _p(i32), _p(i64), _stream()),
               'regtr_ball_query')
    _count(1)
    return i32, i64


# ------------------------------------------------------------------------ encoder

def kpconv(q_pts, s_pts, idx32, x, weights, kernel_points, extent: float, out=None, nq_dev=None, ns_dev=None,
           row_flags=None, instats=None):
    """KPConv.forward (rigid / linear / sum).  idx32 (Nq,K) int32, x (Ns,Cin) -> (Nq,Cout).
    nq_dev / ns_dev: optional 1-element int32 device tensors with the actual counts when the
    leading dimensions are capacities."""
    L = _lib.load()
    _chk(q_pts, torch.float32, 'q_pts', 2); _chk(s_pts, torch.float32, 's_pts', 2)
    _chk(idx32, torch.int32, 'neighb_inds', 2); _chk(x, torch.float32, 'x', 2)
    _chk(weights, torch.float32, 'weights', 3); _chk(kernel_points, torch.float32, 'kernel_points', 2)
    Nq, K = idx32.shape
    Ns, Cin = x.shape
    P, Cin_w, Cout = weights.shape
    if P != 15 or kernel_points.shape != (15, 3) or Cin_w != Cin or q_pts.shape[0] != Nq or s_pts.shape[0] != Ns:
        raise ValueError('kpconv: inconsistent shapes')
    out = torch.empty((Nq, Cout), dtype=torch.float32, device=x.device) if out is None else out
    stats = None
    if instats is not None and (Cin == 1 or Cout % 32):
        raise ValueError('kpconv: the statistics epilogue needs Cin > 1 and Cout % 32 == 0')
    nb = L.regtr_kpconv_fwd_ws_bytes(Nq, Ns, Cin, Cout) if Cin == 1 else L.regtr_kpconv_ws_bytes(Nq, Ns, Cin)
    ws = workspace(nb, x.device, 'kpconv')
    trace = KPCONV_TRACE
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if trace is not None else None
    if ev:
        ev[0].record()
    if Cin == 1:
        # first block: gather + aggregation + the 15 x Cout contraction in one kernel (k_kpconv_c1)
        _lib.check(L.regtr_kpconv_fwd(_p(q_pts), _p(s_pts), _p(idx32), _p(x), _p(weights), _p(kernel_points), Nq, Ns,
                                      _p(nq_dev), _p(ns_dev), K, Cin, Cout, float(extent), _p(out), _p(ws),
                                      ws.numel(), _stream()), 'regtr_kpconv_fwd')
        _count(1)
    else:
        # gather/aggregate kernel, then the [Nq,15Cin] x [15Cin,Cout] contraction on the tensor cores
        if (15 * Cin) % 4:
            raise _lib.RegtrLibError(f'kpconv: Cin={Cin} breaks the 16-byte row pitch of the contraction (no fallback)')
        wf = ws[:Nq * 15 * Cin * 4].view(torch.float32).view(Nq, 15 * Cin)
        flags = row_flags if row_flags is not None else ws[regtr_align_up(Nq * 15 * Cin * 4):]
        _lib.check(L.regtr_kpconv_aggregate(_p(q_pts), _p(s_pts), _p(idx32), _p(x), _p(kernel_points), Nq, Ns,
                                            _p(nq_dev), _p(ns_dev), K, Cin, float(extent), _p(wf), _p(flags),
                                            1 if row_flags is not None else 0, _stream()), 'regtr_kpconv_aggregate')
        _count(1 if row_flags is not None else 2)
        hi, lo = split_weight(weights.view(15 * Cin, Cout), transpose=True)
        if ev:
            ev[1].record()
            ev.append(True)
        if instats is not None:                 # (offs, n_clouds): statistics of the output in the GEMM epilogue
            out, stats = gemm_instats(wf, hi, lo, instats[0], instats[1], m_dev=nq_dev, out=out)
        else:
            gemm(wf, hi, lo, m_dev=nq_dev, out=out)
    if ev:
        ev[2].record()
        trace.append((ev[0], ev[2], dict(Nq=Nq, Ns=Ns, K=K, Cin=Cin, Cout=Cout, idx=idx32,
                                         mid=ev[1] if len(ev) > 3 else None,
                                         args=(q_pts, s_pts, idx32, x, kernel_points, float(extent)),
                                         row_flags=row_flags)))
    return out if instats is None else (out, stats)


def kpconv_aggregate(q_pts, s_pts, idx32, x, kernel_points, extent: float, wf=None, nq_dev=None, ns_dev=None,
                     row_flags=None):
    """Gather + influence + aggregation only: -> wf (Nq, 15*Cin) (already / neighbour count)."""
    L = _lib.load()
    Nq, K = idx32.shape
    Ns, Cin = x.shape
    wf = torch.empty((Nq, 15 * Cin), dtype=torch.float32, device=x.device) if wf is None else wf
    flags = row_flags if row_flags is not None else workspace(max(Ns, 1), x.device, 'rowflags')
    _lib.check(L.regtr_kpconv_aggregate(_p(q_pts), _p(s_pts), _p(idx32), _p(x), _p(kernel_points), Nq, Ns,
                                        _p(nq_dev), _p(ns_dev), K, Cin, float(extent), _p(wf), _p(flags),
                                        1 if row_flags is not None else 0, _stream()), 'regtr_kpconv_aggregate')
    _count(1 if (row_flags is not None or Cin == 1) else 2)
    return wf


def max_pool(x, idx32, ns_dev=None):
    L = _lib.load()
    _chk(x, torch.float32, 'x', 2); _chk(idx32, torch.int32, 'inds', 2)
    Nq, K = idx32.shape
    out = torch.empty((Nq, x.shape[1]), dtype=torch.float32, device=x.device)
    _lib.check(L.regtr_max_pool(_p(x), _p(idx32), Nq, x.shape[0], _p(ns_dev), K, x.shape[1], _p(out), _stream()),
               'regtr_max_pool')
    _count(1)
    return out


def instnorm_act(x, offs, n_clouds: int, res=None, slope: float = -1.0, eps: float = 1e-5, out=None,
                 want_flags: bool = False):
    """out = act(InstanceNorm_per_cloud(x) + res); slope < 0 -> no activation.
    want_flags: also return the per-row `sum > 0` flags the consuming KPConv needs (uint8, n rows)."""
    L = _lib.load()
    _chk(x, torch.float32, 'x', 2); _chk(offs, torch.int32, 'offs', 1)
    n, C = x.shape
    if res is not None:
        _chk(res, torch.float32, 'res', 2)
    out = torch.empty_like(x) if out is None else out
    nb = L.regtr_instnorm_ws_bytes(n, n_clouds, C)
    ws = workspace(nb, x.device, 'instnorm')
    flags = None
    if want_flags and C // 4 <= 32 and (C // 4) & (C // 4 - 1) == 0:
        flags = torch.empty(max(n, 1), dtype=torch.uint8, device=x.device)
    # self-resetting completion counters: the last statistics block of a (cloud, channel tile) finalises it
    cnt = workspace(L.regtr_instnorm_counter_bytes(n_clouds, C), x.device, 'instnorm_cnt', zero=True)
    _lib.check(L.regtr_instnorm_act(_p(x), _p(offs), n_clouds, n, C, float(eps), _p(res), float(slope), _p(out),
                                    _p(flags), _p(ws), ws.numel(), _p(cnt), _stream()), 'regtr_instnorm_act')
    _count(2)
    return (out, flags) if want_flags else out


def instnorm_apply(x, offs, n_clouds: int, stats, res=None, slope: float = -1.0, out=None, want_flags: bool = False):
    """Apply pass of the per-cloud InstanceNorm with statistics from `gemm_instats`:
    out = act((x - mean) * rstd + res)."""
    L = _lib.load()
    _chk(x, torch.float32, 'x', 2); _chk(offs, torch.int32, 'offs', 1); _chk(stats, torch.float32, 'stats', 3)
    n, C = x.shape
    out = torch.empty_like(x) if out is None else out
    flags = None
    if want_flags and C // 4 <= 32 and (C // 4) & (C // 4 - 1) == 0:
        flags = torch.empty(max(n, 1), dtype=torch.uint8, device=x.device)
    _lib.check(L.regtr_instnorm_apply(_p(x), _p(offs), n_clouds, n, C, _p(stats), _p(res), float(slope), _p(out),
                                      _p(flags), _stream()), 'regtr_instnorm_apply')
    _count(1)
    return (out, flags) if want_flags else out


# -------------------------------------------------------------------- dense layers

def split_weight(w: torch.Tensor, transpose: bool = False):
    """(hi, lo) TF32 halves of a weight matrix [N,K] (or of its transpose).  Cached ON the owning
    parameter object (so the cache dies with the model and can never alias a recycled address),
    keyed by view geometry and the parameter's version counter: inference pays the split once."""
    L = _lib.load()
    owner = w._base if w._base is not None else w
    cache = owner.__dict__.setdefault('_regtr_split', {})
    key = (w.storage_offset(), tuple(w.shape), tuple(w.stride()), transpose, owner._version)
    hit = cache.get(key)
    if hit is not None:
        return hit
    src = (w.detach().t() if transpose else w.detach()).contiguous().to(torch.float32)
    hi, lo = torch.empty_like(src), torch.empty_like(src)
    _lib.check(L.regtr_split_tf32(_p(src), src.numel(), _p(hi), _p(lo), _stream()), 'regtr_split_tf32')
    _count(1)
    for k in [k for k in cache if k[-1] != owner._version]:
        del cache[k]
    cache[key] = (hi, lo)
    return hi, lo


def gemm(a, b_hi, b_lo, bias=None, residual=None, relu=False, m_dev=None, out=None):
    """act(a @ B^T + bias + residual) with B = b_hi + b_lo ([N,K] row-major), 3xTF32 tcgen05 kernel."""
    L = _lib.load()
    if not a.is_cuda or a.dtype != torch.float32 or a.dim() != 2 or a.stride(1) != 1:
        raise ValueError('gemm: A must be a CUDA fp32 matrix with unit column stride')
    M, K = a.shape
    N = b_hi.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device) if out is None else out
    nb = L.regtr_gemm_ws_bytes(M, N, K)
    ws = workspace(nb, a.device, 'gemm')
    if TRACE is not None:
        TRACE.append(('gemm', dict(M=M, N=N, K=K, split_k=nb > 256),
                      lambda: gemm(a, b_hi, b_lo, bias=bias, residual=residual, relu=relu, m_dev=m_dev, out=out)))
    _lib.check(L.regtr_gemm_tf32x3(_p(a), a.stride(0), _p(b_hi), _p(b_lo), b_hi.stride(0), _p(out), out.stride(0),
                                   _p(bias), _p(residual), residual.stride(0) if residual is not None else 0,
                                   M, N, K, _p(m_dev), 1 if relu else 0, _p(ws), ws.numel(), _stream()),
               'regtr_gemm_tf32x3')
    _count(2 if nb > 256 else 1)
    return out


def gemm_instats(a, b_hi, b_lo, offs, n_clouds: int, eps: float = 1e-5, m_dev=None, out=None):
    """a @ B^T plus the per-cloud InstanceNorm statistics of the result: 32-row partial sums from the GEMM
    epilogue + a small fixed-order finalisation (no atomics).  -> (out (M,N), stats (n_clouds, N, 2) = (mean, rstd))."""
    L = _lib.load()
    if not a.is_cuda or a.dtype != torch.float32 or a.dim() != 2 or a.stride(1) != 1:
        raise ValueError('gemm: A must be a CUDA fp32 matrix with unit column stride')
    M, K = a.shape
    N = b_hi.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device) if out is None else out
    stats = torch.empty((n_clouds, N, 2), dtype=torch.float32, device=a.device)
    nb = L.regtr_gemm_ws_bytes(M, N, K)
    ws = workspace(nb, a.device, 'gemm')
    acc = workspace(L.regtr_instnorm_part_bytes(M, N), a.device, 'instnorm_part')
    if TRACE is not None:
        TRACE.append(('gemm', dict(M=M, N=N, K=K, split_k=nb > 256, instats=True),
                      lambda: gemm_instats(a, b_hi, b_lo, offs, n_clouds, eps, m_dev=m_dev, out=out)))
    _lib.check(L.regtr_gemm_tf32x3_instats(_p(a), a.stride(0), _p(b_hi), _p(b_lo), b_hi.stride(0), _p(out), out.stride(0),
                                           M, N, K, _p(m_dev), _p(offs), n_clouds, float(eps), _p(acc), _p(stats),
                                           _p(ws), ws.numel(), _stream()), 'regtr_gemm_tf32x3_instats')
    _count(3 if nb > 256 else 2)
    return out, stats


def linear_instats(x, weight, offs, n_clouds: int, eps: float = 1e-5, m_dev=None):
    """nn.Linear(bias=False) followed by InstanceNorm statistics (UnaryBlock, kpconv_blocks.py:546-561)."""
    if x.shape[1] % 4 or x.stride(0) % 4 or x.data_ptr() % 16:
        raise _lib.RegtrLibError(f'linear: K={x.shape[1]}: rows must be 16-byte aligned multiples of 4 floats')
    hi, lo = split_weight(weight)
    return gemm_instats(x, hi, lo, offs, n_clouds, eps, m_dev=m_dev)


def linear(x, weight, bias=None, residual=None, relu=False, m_dev=None):
    """nn.Linear forward (x @ weight^T + bias) (+ residual, + ReLU) on the 3xTF32 tcgen05 GEMM of this
    library.  The TMA row pitch needs K % 4 == 0 and 16-byte aligned rows; anything else raises (callers
    with an odd K zero-pad it, see PositionEmbeddingLearned) -- there is no library fallback."""
    if x.shape[1] % 4 or x.stride(0) % 4 or x.data_ptr() % 16:
        raise _lib.RegtrLibError(f'linear: K={x.shape[1]}, row stride {x.stride(0)}: rows must be 16-byte '
                                 'aligned multiples of 4 floats (zero-pad K); no cuBLAS fallback')
    hi, lo = split_weight(weight)
    return gemm(x, hi, lo, bias=bias, residual=residual, relu=relu, m_dev=m_dev)


# -------------------------------------------------------------------- transformer

_dim_t_cache = {}


def sine_dim_t(n_freq: int, temperature: float, device):
    """The reference's frequency table, computed with the same torch fp32 ops
    (position_embedding.py:39-40) and cached per device."""
    key = (n_freq, float(temperature), str(device))
    if key not in _dim_t_cache:
        d = torch.arange(n_freq, dtype=torch.float32)
        d = temperature ** (2 * torch.div(d, 2, rounding_mode='trunc') / n_freq)
        _dim_t_cache[key] = d.to(device)
    return _dim_t_cache[key]


def pos_embed_sine(xyz, d_model: int = 256, temperature: float = 10000.0, scale: float = 1.0):
    L = _lib.load()
    _chk(xyz, torch.float32, 'xyz', 2)
    n, n_dim = xyz.shape
    if n_dim != 3:
        raise ValueError('pos_embed_sine: only 3-D coordinates are on the hot path')
    n_freq = d_model // n_dim // 2 * 2
    out = torch.empty((n, d_model), dtype=torch.float32, device=xyz.device)
    s32 = torch.tensor(scale * 2 * math.pi, dtype=torch.float32).item()
    _lib.check(L.regtr_pos_embed_sine(_p(xyz), n, _p(sine_dim_t(n_freq, temperature, xyz.device)), n_freq, d_model,
                                      s32, _p(out), _stream()), 'regtr_pos_embed_sine')
    _count(1)
    return out


def layernorm_pos(x, gamma, beta, pos=None, eps: float = 1e-5, want_plain=True, want_pos=True, n_dev=None):
    """-> (LN(x), LN(x)+pos); either may be skipped.  n_dev: device row count when x is capacity-shaped."""
    L = _lib.load()
    _chk(x, torch.float32, 'x', 2)
    n, E = x.shape
    y = torch.empty_like(x) if want_plain else None
    yp = torch.empty_like(x) if want_pos else None
    _lib.check(L.regtr_layernorm_pos(_p(x), _p(gamma), _p(beta), _p(pos), n, _p(n_dev), E, float(eps), _p(y), _p(yp),
                                     _stream()), 'regtr_layernorm_pos')
    _count(1)
    return y, yp


def attention_plan(offs, B: int):
    """Device-side (6, 2B + 1) int32 table: q_start, q_len, cross k_start, cross k_len, then the exclusive prefixes
    of the 64- and 128-query tile counts per problem (entry 2B = total).  No host sync."""
    L = _lib.load()
    _chk(offs, torch.int32, 'offs', 1)
    plan = torch.empty((6, 2 * B + 1), dtype=torch.int32, device=offs.device)
    _lib.check(L.regtr_attention_plan(_p(offs), B, _p(plan), _stream()), 'regtr_attention_plan')
    _count(1)
    return plan


def corr_decode(qp, kp, xyz, q_start, q_len, k_start, k_len, max_q_len: int, n_layers: int, out=None):
    """CorrespondenceDecoder.simple_attention for all decoder layers: qp/kp (n_layers*N, D) projected
    queries / keys, xyz (N,3) -> (n_layers*N, 3) attention-weighted key coordinates."""
    L = _lib.load()
    _chk(qp, torch.float32, 'qp', 2); _chk(kp, torch.float32, 'kp', 2); _chk(xyz, torch.float32, 'xyz', 2)
    rows, D = qp.shape
    N = xyz.shape[0]
    if kp.shape != qp.shape or rows != n_layers * N or qp.stride(0) != kp.stride(0) or qp.stride(1) != 1:
        raise ValueError('corr_decode: inconsistent shapes')
    out = torch.zeros((rows, 3), dtype=torch.float32, device=qp.device) if out is None else out
    _lib.check(L.regtr_corr_decode_fwd(_p(qp), _p(kp), qp.stride(0), _p(xyz.contiguous()), _p(out), _p(q_start),
                                       _p(q_len), _p(k_start), _p(k_len), int(q_start.numel()), int(max_q_len),
                                       int(n_layers), N, D, 1.0 / math.sqrt(D), _stream()), 'regtr_corr_decode_fwd')
    _count(1)
    return out


def mha_varlen(q, k, v, q_start, q_len, k_start, k_len, max_q_len: int, n_heads: int, out=None, tiles=None):
    """softmax(q k^T / sqrt(dh)) v per head over explicit (query range, key range) problems.
    q/k/v may be column slices of a wider row-major matrix (stride(0) is the leading dim)."""
    L = _lib.load()
    for t, nm in ((q, 'q'), (k, 'k'), (v, 'v')):
        if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
            raise ValueError(f'mha_varlen: {nm} must be a CUDA fp32 matrix with unit column stride')
    E = q.shape[1]
    dh = E // n_heads
    # rows outside every problem (capacity padding) stay uninitialised: every consumer is row-wise and skips them
    out = torch.empty((q.shape[0], E), dtype=torch.float32, device=q.device) if out is None else out
    if TRACE is not None:
        ql, kl = q_len.tolist(), k_len.tolist()
        TRACE.append(('mha', dict(pairs_qk=sum(a * b for a, b in zip(ql, kl)), E=E, tokens=sum(ql)),
                      lambda: mha_varlen(q, k, v, q_start, q_len, k_start, k_len, max_q_len, n_heads, out=out, tiles=tiles)))
    tb, mt = (tiles[0], int(tiles[1])) if tiles is not None else (None, 0)     # (device tile table, host bound)
    _lib.check(L.regtr_mha_varlen_fwd(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out),
                                      out.stride(0), _p(q_start), _p(q_len), _p(k_start), _p(k_len),
                                      q_start.numel(), int(max_q_len), _p(tb), mt, n_heads, dh, 1.0 / math.sqrt(dh),
                                      _stream()), 'regtr_mha_varlen_fwd')
    _count(1)
    return out


def mha_bf16_tc(x, in_w, in_b, q_start, q_len, k_start, k_len, max_q_len: int, n_heads: int, m_dev=None):
    """Attention block core in the fast precision mode: packed in-projection (3xTF32 GEMM, bf16
    epilogue) + tcgen05 bf16 attention.  x (N,E) fp32 (already LN + pos); returns O (N,E) fp32."""
    L = _lib.load()
    _chk(x, torch.float32, 'x', 2)
    N, E = x.shape
    dh = E // n_heads
    hi, lo = split_weight(in_w)
    ld_vt = (N + 63) // 64 * 64 + 64            # token pitch of the transposed V (TMA: 16-byte multiple)
    qk = torch.empty((N, 2 * E), dtype=torch.bfloat16, device=x.device)
    vt = torch.zeros((E, ld_vt), dtype=torch.bfloat16, device=x.device)
    _lib.check(L.regtr_gemm_tf32x3_qkv_bf16(_p(x), x.stride(0), _p(hi), _p(lo), hi.stride(0), _p(in_b), N, 3 * E, E,
                                            2 * E, _p(qk), 2 * E, _p(vt), ld_vt, _p(m_dev), _stream()),
               'regtr_gemm_tf32x3_qkv_bf16')
    out = torch.zeros((N, E), dtype=torch.float32, device=x.device)
    _lib.check(L.regtr_mha_bf16_tc_fwd(_p(qk), 2 * E, _p(vt), ld_vt, N, _p(out), E, _p(q_start), _p(q_len),
                                       _p(k_start), _p(k_len), q_start.numel(), int(max_q_len), n_heads, dh,
                                       1.0 / math.sqrt(dh), _stream()), 'regtr_mha_bf16_tc_fwd')
    _count(2)
    return out


def mha_tf32_tc(x, in_w, in_b, q_start, q_len, k_start, k_len, max_q_len: int, n_heads: int, m_dev=None, tiles=None):
    """Attention block core, fp32-accurate, on the tcgen05 tensor cores: packed in-projection (3xTF32 GEMM whose
    epilogue writes q / k / v^T as TF32 (hi, lo) halves) + the TMA-fed 3xTF32 attention kernel (P in tensor memory).
    x (N,E) fp32 (already LN + pos); returns O (N,E) fp32."""
    L = _lib.load()
    _chk(x, torch.float32, 'x', 2)
    N, E = x.shape
    dh = E // n_heads
    hi, lo = split_weight(in_w)
    ld_vt = (N + 63) // 64 * 64 + 64            # token pitch of the transposed V (16-byte multiple, tile over-read)
    qk4 = torch.empty((N, 4 * E), dtype=torch.float32, device=x.device)
    vt2 = torch.zeros((2 * E, ld_vt), dtype=torch.float32, device=x.device)     # padding tokens must stay finite
    qscale = (1.0 / math.sqrt(dh)) * 1.4426950408889634
    _lib.check(L.regtr_gemm_tf32x3_qkv_split(_p(x), x.stride(0), _p(hi), _p(lo), hi.stride(0), _p(in_b), N, 3 * E, E, E,
                                             float(qscale), _p(qk4), 4 * E, _p(vt2), ld_vt, _p(m_dev), _stream()),
               'regtr_gemm_tf32x3_qkv_split')
    out = torch.empty((N, E), dtype=torch.float32, device=x.device)
    tb, mt = (tiles[0], int(tiles[1])) if tiles is not None else (None, 0)
    if TRACE is not None:
        ql, kl = q_len.tolist(), k_len.tolist()
        TRACE.append(('gemm', dict(M=N, N=3 * E, K=E, split_k=False, qkv_split=True), lambda: L.regtr_gemm_tf32x3_qkv_split(
            _p(x), x.stride(0), _p(hi), _p(lo), hi.stride(0), _p(in_b), N, 3 * E, E, E, float(qscale), _p(qk4), 4 * E,
            _p(vt2), ld_vt, _p(m_dev), _stream())))
        TRACE.append(('mha', dict(pairs_qk=sum(a * b for a, b in zip(ql, kl)), E=E, tokens=sum(ql)),
                      lambda: L.regtr_mha_tf32_tc_fwd(_p(qk4), 4 * E, _p(vt2), ld_vt, N, _p(out), E, _p(q_start), _p(q_len),
                                                      _p(k_start), _p(k_len), q_start.numel(), int(max_q_len), _p(tb), mt,
                                                      n_heads, dh, _stream())))
    _lib.check(L.regtr_mha_tf32_tc_fwd(_p(qk4), 4 * E, _p(vt2), ld_vt, N, _p(out), E, _p(q_start), _p(q_len),
                                       _p(k_start), _p(k_len), q_start.numel(), int(max_q_len), _p(tb), mt, n_heads, dh,
                                       _stream()), 'regtr_mha_tf32_tc_fwd')
    _count(2)
    return out


# --------------------------------------------------------------------------- pose

def kabsch(a, b, w, offs):
    """Packed problems: rows [offs[i], offs[i+1]) -> T (n_problems,3,4)."""
    L = _lib.load()
    _chk(a, torch.float32, 'a', 2); _chk(b, torch.float32, 'b', 2); _chk(w, torch.float32, 'w', 1)
    n_prob = offs.numel() - 1
    T = torch.empty((n_prob, 3, 4), dtype=torch.float32, device=a.device)
    _lib.check(L.regtr_kabsch_fwd(_p(a), _p(b), _p(w), _p(offs), n_prob, _p(T), _stream()), 'regtr_kabsch_fwd')
    _count(1)
    return T


def pose_from_corr(kp, corr, logit, offs, B: int):
    """kp (n,3), corr (L,n,3), logit (L,n), offs (2B+1) -> pose (L,B,3,4)."""
    L_ = _lib.load()
    _chk(kp, torch.float32, 'kp', 2); _chk(corr, torch.float32, 'corr', 3); _chk(logit, torch.float32, 'logit', 2)
    nl, n = logit.shape
    pose = torch.empty((nl, B, 3, 4), dtype=torch.float32, device=kp.device)
    _lib.check(L_.regtr_pose_from_corr(_p(kp), _p(corr), _p(logit), _p(offs), n, B, nl, _p(pose), _stream()),
               'regtr_pose_from_corr')
    _count(1)
    return pose
