"""Deterministic random `state_dict` in the reference's checkpoint layout.

No pretrained RegTR weights are available offline, so parity
tests, goldens and the benchmark all share seeded random weights generated here
and copied by `load_state_dict` into whichever implementation is under test.
The key set and shapes are the reference's 168-key layout (SURVEY.md 8b;
reference constructors: src/models/regtr.py:23-102,
src/models/backbone_kpconv/kpconv_blocks.py:590-704,
src/models/transformer/transformers.py:84-119).

Kernel points: the reference loads a 15-point disposition file, applies a random
z-rotation and N(0, 0.01) noise and scales by the conv radius
(src/models/backbone_kpconv/kernels/kernel_points.py:387-469); a checkpoint then
carries the result.  `kernel_disposition` builds a geometrically equivalent
disposition from scratch (centre + 14 quasi-uniform shell points at 0.66 r) so
the package needs no data file; real checkpoints overwrite it on load.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from .config import pyramid_plan


def kernel_disposition(radius: float, num_kpoints: int = 15, rng=None) -> np.ndarray:
    """(num_kpoints,3) float32: centre point + Fibonacci-sphere shell, rotated about z, jittered."""
    rng = rng or np.random.default_rng(0)
    n = num_kpoints - 1
    k = np.arange(n) + 0.5
    z = 1.0 - 2.0 * k / n
    phi = k * math.pi * (3.0 - math.sqrt(5.0))
    rho = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    shell = 0.661 * np.stack([rho * np.cos(phi), rho * np.sin(phi), z], axis=1)
    pts = np.concatenate([np.zeros((1, 3)), shell], axis=0)
    theta = rng.random() * 2.0 * math.pi
    c, s = math.cos(theta), math.sin(theta)
    R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    pts = pts + rng.normal(scale=0.01, size=pts.shape)
    return ((radius * pts) @ R).astype(np.float32)


def state_dict_spec(cfg):
    """Ordered {key: shape} of the reference checkpoint for `cfg` (parameters only)."""
    _, blocks, enc_out = pyramid_plan(cfg)
    P = int(cfg.num_kernel_points)
    spec = OrderedDict()
    for i, b in enumerate(blocks):
        pre = f'kpf_encoder.encoder_blocks.{i}.'
        if b['kind'] == 'simple':
            spec[pre + 'KPConv.weights'] = (P, b['in_dim'], b['out_dim'] // 2)
            spec[pre + 'KPConv.kernel_points'] = (P, 3)
            continue
        mid = b['out_dim'] // 4
        if b['in_dim'] != mid:
            spec[pre + 'unary1.mlp.weight'] = (mid, b['in_dim'])
        spec[pre + 'KPConv.weights'] = (P, mid, mid)
        spec[pre + 'KPConv.kernel_points'] = (P, 3)
        spec[pre + 'unary2.mlp.weight'] = (b['out_dim'], mid)
        if b['in_dim'] != b['out_dim']:
            spec[pre + 'unary_shortcut.mlp.weight'] = (b['out_dim'], b['in_dim'])
    E, Fd = int(cfg.d_embed), int(cfg.d_feedforward)
    spec['feat_proj.weight'] = (E, enc_out)
    spec['feat_proj.bias'] = (E,)
    for i in range(int(cfg.num_encoder_layers)):
        pre = f'transformer_encoder.layers.{i}.'
        for m in ('self_attn', 'multihead_attn'):
            spec[pre + m + '.in_proj_weight'] = (3 * E, E)
            spec[pre + m + '.in_proj_bias'] = (3 * E,)
            spec[pre + m + '.out_proj.weight'] = (E, E)
            spec[pre + m + '.out_proj.bias'] = (E,)
        spec[pre + 'linear1.weight'] = (Fd, E)
        spec[pre + 'linear1.bias'] = (Fd,)
        spec[pre + 'linear2.weight'] = (E, Fd)
        spec[pre + 'linear2.bias'] = (E,)
        for n in ('norm1', 'norm2', 'norm3'):
            spec[pre + n + '.weight'] = (E,)
            spec[pre + n + '.bias'] = (E,)
    if cfg.pre_norm:                                      # regtr.py:64: final norm only for pre-norm stacks
        spec['transformer_encoder.norm.weight'] = (E,)
        spec['transformer_encoder.norm.bias'] = (E,)
    learned = cfg.get('pos_emb_type', 'sine') == 'learned'
    pe_dims = [(32, 3), (64, 32), (128, 64), (256, 128), (E, 256)]   # position_embedding.py:59-69
    if learned:
        for i, (o, k) in zip((0, 2, 4, 6, 8), pe_dims):
            spec[f'pos_embed.mlp.{i}.weight'] = (o, k)
            spec[f'pos_embed.mlp.{i}.bias'] = (o,)
    if cfg.get('direct_regress_coor', False):
        pre = 'correspondence_decoder.'
        spec[pre + 'coor_mlp.0.weight'] = (E, E); spec[pre + 'coor_mlp.0.bias'] = (E,)
        spec[pre + 'coor_mlp.2.weight'] = (E, E); spec[pre + 'coor_mlp.2.bias'] = (E,)
        spec[pre + 'coor_mlp.4.weight'] = (3, E); spec[pre + 'coor_mlp.4.bias'] = (3,)
        spec[pre + 'conf_logits_decoder.weight'] = (1, E)
        spec[pre + 'conf_logits_decoder.bias'] = (1,)
    else:                                                 # CorrespondenceDecoder (regtr.py:298-311)
        pre = 'correspondence_decoder.'
        if learned:                                       # the shared embedding module is registered twice
            for i, (o, k) in zip((0, 2, 4, 6, 8), pe_dims):
                spec[pre + f'pos_embed.mlp.{i}.weight'] = (o, k)
                spec[pre + f'pos_embed.mlp.{i}.bias'] = (o,)
        spec[pre + 'q_norm.weight'] = (E,); spec[pre + 'q_norm.bias'] = (E,)
        spec[pre + 'q_proj.weight'] = (E, E); spec[pre + 'q_proj.bias'] = (E,)
        spec[pre + 'k_proj.weight'] = (E, E); spec[pre + 'k_proj.bias'] = (E,)
        spec[pre + 'conf_logits_decoder.weight'] = (1, E)
        spec[pre + 'conf_logits_decoder.bias'] = (1,)
    if cfg.feature_loss_type == 'infonce':
        spec['feature_criterion.W'] = (E, E)
        spec['feature_criterion_un.W'] = (E, E)
    return spec


def random_state_dict(cfg, seed: int = 0, spread_corr: bool = True):
    """Seeded fp32 CPU state_dict.  Scales follow the reference initialisers' fan-in rule.

    `spread_corr` scales the last regressor layer so that predicted correspondences are
    spread over ~1 m instead of collapsing to a point: with collapsed correspondences the
    Kabsch covariance is near-singular and the pose amplifies 1e-6 feature noise to 1e-5
    (SURVEY.md H4), which would test conditioning rather than the kernels.
    """
    rng = np.random.default_rng(seed)
    _, blocks, _ = pyramid_plan(cfg)
    sd = OrderedDict()
    for key, shape in state_dict_spec(cfg).items():
        if key.endswith('kernel_points'):
            bi = int(key.split('.')[2])
            arr = kernel_disposition(blocks[bi]['radius'], shape[0], rng)
        elif key.endswith('KPConv.weights'):
            bound = 1.0 / math.sqrt(shape[1] * shape[2])
            arr = rng.uniform(-bound, bound, size=shape)
        elif ('.norm' in key or 'q_norm' in key) and key.endswith('weight'):
            arr = 1.0 + 0.1 * rng.standard_normal(shape)
        elif ('.norm' in key or 'q_norm' in key) and key.endswith('bias'):
            arr = 0.1 * rng.standard_normal(shape)
        elif key.endswith('.W'):
            arr = np.eye(shape[0])
        elif key.endswith('bias'):
            arr = rng.uniform(-0.05, 0.05, size=shape)
        else:
            bound = 1.0 / math.sqrt(shape[-1])
            arr = rng.uniform(-bound, bound, size=shape)
            if spread_corr and key.endswith('coor_mlp.4.weight'):
                arr = arr * 8.0
            if spread_corr and key.endswith('conf_logits_decoder.weight'):
                arr = arr * 4.0
        sd[key] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
    for key in list(sd):                                  # one module, two names: same tensors
        if key.startswith('correspondence_decoder.pos_embed.'):
            sd[key] = sd[key[len('correspondence_decoder.'):]]
    return sd
