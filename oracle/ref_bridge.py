"""ORACLE (test infrastructure only) -- import the UNMODIFIED reference on the CPU.

Only usable where /root/reference exists (the build container); nothing in the
`-m gpu` tests, smoke() or bench.py depends on it.  It is how the restatement in
oracle/regtr_oracle.py is pinned: tests/golden/make_golden.py runs the reference
modules through this bridge and commits the outputs as fixtures.

Recipe: the reference imports MinkowskiEngine,
pytorch3d, open3d, vtk, nibabel, matplotlib, h5py, coloredlogs, easydict at
module scope; none is installed.  The unused ones are replaced by MagicMock; the
two that ARE used on the hot path get functional stand-ins built on the oracle's
deterministic restatements (oracle/pre.py), i.e. exactly the "ORACLE-G"
definition of SURVEY.md 8c.  The reference source files are not modified or
copied.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

REF_SRC = '/root/reference/src'


def available() -> bool:
    return os.path.isdir(REF_SRC)


def _install_stubs():
    from . import pre

    for name in ('nibabel', 'nibabel.quaternions', 'matplotlib', 'matplotlib.pyplot',
                 'matplotlib.colors', 'open3d', 'vtk', 'vtk.util', 'vtk.util.numpy_support',
                 'h5py', 'coloredlogs', 'tensorboardX', 'torch.utils.tensorboard'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mock.MagicMock()

    if 'easydict' not in sys.modules:
        ed = types.ModuleType('easydict')
        from regtr_b200.config import Cfg
        ed.EasyDict = Cfg
        sys.modules['easydict'] = ed

    # ---- pytorch3d.ops: packed_to_padded + ball_query (first-K in index order, -1 padded)
    p3d = types.ModuleType('pytorch3d')
    ops = types.ModuleType('pytorch3d.ops')

    def packed_to_padded(x, first_idx, max_size):
        B = first_idx.shape[0]
        out = x.new_zeros((B, int(max_size)) + tuple(x.shape[1:]))
        ends = list(first_idx[1:].tolist()) + [x.shape[0]]
        for b in range(B):
            a, e = int(first_idx[b]), int(ends[b])
            out[b, :e - a] = x[a:e]
        return out

    def ball_query(p1, p2, lengths1, lengths2, K, radius):
        B = p1.shape[0]
        idx = torch.full((B, p1.shape[1], K), -1, dtype=torch.int64)
        for b in range(B):
            n1, n2 = int(lengths1[b]), int(lengths2[b])
            r = pre.ball_query(p1[b, :n1].numpy(), [n1], p2[b, :n2].numpy(), [n2], K, radius)
            r[r >= n2] = -1
            idx[b, :n1] = torch.from_numpy(r)
        return types.SimpleNamespace(idx=idx, dists=None, knn=None)

    ops.packed_to_padded = packed_to_padded
    ops.ball_query = ball_query
    p3d.ops = ops
    sys.modules['pytorch3d'] = p3d
    sys.modules['pytorch3d.ops'] = ops

    # ---- MinkowskiEngine: batched_coordinates + SparseTensor(UNWEIGHTED_AVERAGE)
    me = types.ModuleType('MinkowskiEngine')
    me.utils = types.SimpleNamespace()

    def batched_coordinates(coords, device=None):
        rows = []
        for b, c in enumerate(coords):
            c = torch.floor(c).to(torch.int32)
            rows.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=torch.int32), c], 1))
        return torch.cat(rows, 0)

    class _QMode:
        UNWEIGHTED_AVERAGE = 'unweighted_average'

    class SparseTensor:
        """Voxel mean in canonical (batch, x, y, z) order with index-ordered fp32 sums."""

        def __init__(self, features, coordinates, quantization_mode=None):
            assert quantization_mode == _QMode.UNWEIGHTED_AVERAGE
            c = coordinates.numpy().astype(np.int64)
            f = features.numpy().astype(np.float32)
            order = np.lexsort((np.arange(len(c)), c[:, 3], c[:, 2], c[:, 1], c[:, 0]))
            cs = c[order]
            new = np.ones(len(cs), dtype=bool)
            new[1:] = np.any(cs[1:] != cs[:-1], axis=1)
            starts = np.nonzero(new)[0]
            ends = np.append(starts[1:], len(cs))
            out = np.empty((len(starts), f.shape[1]), dtype=np.float32)
            for m, (a, e) in enumerate(zip(starts, ends)):
                acc = np.zeros(f.shape[1], dtype=np.float32)
                for i in order[a:e]:
                    acc = acc + f[i]
                out[m] = acc / np.float32(e - a)
            self.features = torch.from_numpy(out)
            batch = cs[starts, 0]
            nb = int(c[:, 0].max()) + 1 if len(c) else 0
            self.decomposed_features = [self.features[torch.from_numpy(batch == b)] for b in range(nb)]

    me.utils.batched_coordinates = batched_coordinates
    me.SparseTensorQuantizationMode = _QMode
    me.SparseTensor = SparseTensor
    sys.modules['MinkowskiEngine'] = me


@contextlib.contextmanager
def _in_ref_dir():
    cwd = os.getcwd()
    os.chdir(REF_SRC)          # kernel_points.py:390 uses a cwd-relative directory
    try:
        yield
    finally:
        os.chdir(cwd)


_MODULES = None


def modules():
    """Import and return the reference modules (models.regtr, kpconv, kpconv_blocks, ...)."""
    global _MODULES
    if _MODULES is None:
        if not available():
            raise RuntimeError('/root/reference is not present on this machine')
        _install_stubs()
        if REF_SRC not in sys.path:
            sys.path.insert(0, REF_SRC)
        with _in_ref_dir():
            import models.regtr as regtr
            import models.backbone_kpconv.kpconv as kpconv
            import models.backbone_kpconv.kpconv_blocks as blocks
            import models.transformer.transformers as transformers
            import models.transformer.position_embedding as posemb
            import utils.se3_torch as se3
        _MODULES = types.SimpleNamespace(regtr=regtr, kpconv=kpconv, blocks=blocks,
                                         transformers=transformers, posemb=posemb, se3=se3)
    return _MODULES


def build_reference_model(cfg, state_dict=None):
    """Instantiate the reference RegTR(cfg) on the CPU in eval mode, optionally loading weights."""
    m = modules()
    with _in_ref_dir():
        model = m.regtr.RegTR(cfg.copy())
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=True), None
    return model.eval()


def reference_forward(model, src_list, tgt_list):
    batch = {'src_xyz': [torch.as_tensor(np.asarray(s)) for s in src_list],
             'tgt_xyz': [torch.as_tensor(np.asarray(t)) for t in tgt_list]}
    with torch.no_grad():
        out = model(batch)
    out['kpconv_meta'] = batch['kpconv_meta']
    return out
