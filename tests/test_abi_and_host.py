"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol (no compute
calls), the state_dict layout matches the reference's, config/pyramid plan, and the N>1 sharding
logic over gloo (world_size 2)."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from regtr_b200 import lib
    L = lib.load()
    names = lib.header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f'{n} declared in include/regtr_b200.h but not exported'
        assert n in lib.SIGNATURES, f'{n} has no ctypes signature'
    assert L.regtr_version() == 1
    assert b'sm_100a' in L.regtr_build_info()
    # size queries are host-only and must be callable without a GPU
    assert L.regtr_grid_subsample_ws_bytes(40000) > 40000 * 24
    assert L.regtr_cellgrid_bytes(1000) >= 1000 * 24
    assert L.regtr_kpconv_ws_bytes(100, 100, 32) >= 100 * 15 * 32 * 4


def test_product_path_fails_loudly_without_cuda():
    from regtr_b200 import ops
    from regtr_b200.lib import RegtrLibError
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    x = torch.zeros((4, 32))
    with pytest.raises(RegtrLibError):
        ops.instnorm_act(x, torch.zeros(2, dtype=torch.int32), 1)


def test_state_dict_layout_matches_reference_key_count():
    from regtr_b200.config import get_config
    from regtr_b200.regtr import RegTR
    from regtr_b200.weights import random_state_dict
    for name, n_keys in (('3dmatch', 168), ('modelnet', 146)):       # 168 keys
        cfg = get_config(name)
        m = RegTR(cfg)
        assert len(m.state_dict()) == n_keys
        m.load_state_dict(random_state_dict(cfg, 0), strict=True)
    sd = RegTR(get_config('3dmatch')).state_dict()
    assert sd['kpf_encoder.encoder_blocks.1.KPConv.weights'].shape == (15, 32, 32)
    assert sd['kpf_encoder.encoder_blocks.10.unary2.mlp.weight'].shape == (1024, 256)
    assert sd['transformer_encoder.layers.5.multihead_attn.in_proj_weight'].shape == (768, 256)
    assert sd['feature_criterion.W'].shape == (256, 256)


def test_pyramid_plan_matches_reference_numbers():
    from regtr_b200.config import get_config, pyramid_plan
    levels, blocks, enc_out = pyramid_plan(get_config('3dmatch'))
    assert [round(l['radius'], 6) for l in levels] == [0.0625, 0.125, 0.25, 0.5]
    assert [l['dl'] for l in levels][:3] == [0.05, 0.1, 0.2] and levels[3]['dl'] is None
    assert enc_out == 1024 and len(blocks) == 11
    assert [(b['in_dim'], b['out_dim']) for b in blocks][:4] == [(1, 128), (64, 128), (128, 128), (128, 256)]
    assert abs(blocks[0]['extent'] - 0.05) < 1e-12                   # r * KP_extent / conv_radius


def test_shard_range_partitions():
    from regtr_b200.dist import shard_range
    for n in (1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, n_pairs, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from regtr_b200.dist import gather_poses, shard_range
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    lo, hi = shard_range(n_pairs, rank, world)
    local = torch.stack([torch.full((hi - lo, 3, 4), float(l)) + torch.arange(lo, hi).view(-1, 1, 1) * 10
                         for l in range(6)])
    full = gather_poses(local, n_pairs)
    q.put((rank, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_poses_world2_gloo():
    world, n_pairs = 2, 5
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in procs]
    want = np.stack([np.full((n_pairs, 3, 4), float(l)) + np.arange(n_pairs).reshape(-1, 1, 1) * 10
                     for l in range(6)])
    for r in range(world):
        assert np.array_equal(res[r], want)


def test_level_capacities_are_monotone_and_roomy():
    from regtr_b200.config import get_config
    from regtr_b200.kpconv import level_capacities
    cfg = get_config('3dmatch')
    for cap0 in (8192, 40960, 327680):
        caps = level_capacities(cfg, cap0, ratio=0.30)
        assert len(caps) == 4 and caps[0] == cap0
        assert all(a >= b for a, b in zip(caps, caps[1:]))
        assert all(c % 256 == 0 for c in caps[1:])
        # real 3DMatch keeps 26-27 % of the points per level (38061 -> 10088 -> 2753 -> 751)
        assert caps[1] >= 0.27 * cap0 and caps[3] >= 0.27 ** 3 * cap0
    assert level_capacities(get_config('modelnet'), 8192)[1] <= 8192


def test_header_and_binding_signatures_agree_in_arity():
    """Every C prototype in include/regtr_b200.h has as many parameters as its ctypes signature."""
    import re
    from regtr_b200 import lib
    text = re.sub(r'/\*.*?\*/', '', open(lib.HEADER).read(), flags=re.S)
    for name, (_, args) in lib.SIGNATURES.items():
        m = re.search(r'\b' + name + r'\s*\(([^;]*?)\)\s*;', text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ('', 'void') else params.count(',') + 1
        assert n == len(args), f'{name}: header has {n} parameters, binding has {len(args)}'
