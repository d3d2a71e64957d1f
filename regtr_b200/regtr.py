"""RegTR forward on the B200 kernels -- the drop-in boundary.

`RegTR(cfg)` keeps the reference's constructor, sub-module names, state_dict layout
(168 keys for the 3DMatch config, loadable with strict=True) and
`forward(batch: dict) -> dict` contract (/root/reference/src/models/regtr.py:23-235):

    batch['src_xyz'], batch['tgt_xyz']: lists (B) of (Ni,3) CUDA tensors
    -> outputs: src_feat_un/tgt_feat_un (tuples of (S,256)), src_feat/tgt_feat (lists of
       (6,S,256)), src_kp/tgt_kp, src_kp_warped/tgt_kp_warped (lists of (6,S,3)),
       src_overlap/tgt_overlap (lists of (6,S,1) logits), pose (6,B,3,4);
       side effect batch['kpconv_meta'] (regtr.py:118).

What differs from the reference is *how*: packed tokens instead of padded ones, one host
sync for the pyramid sizes, hand-written sm_100a kernels for the neighbour search, the
KPConv gather/aggregation, normalisations, attention core and Kabsch, and a fused
correspondence-assembly + sigmoid + Kabsch kernel (regtr.py:185-203 in one launch).
Training (`compute_loss`, autograd through the custom kernels) is a "next" row.
"""
from __future__ import annotations

import logging

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .kpconv import KPFEncoder, PreprocessorGPU
from .transformer import (AttentionPlan, PositionEmbeddingCoordsSine, TransformerCrossEncoder,
                          TransformerCrossEncoderLayer)


class _LossParam(nn.Module):
    """Keeps `feature_criterion*.W` loadable (InfoNCELossFull's only parameter; training-only)."""

    def __init__(self, d_embed):
        super().__init__()
        self.W = nn.Parameter(torch.zeros(d_embed, d_embed), requires_grad=True)


class CorrespondenceRegressor(nn.Module):
    """MLP 256->256->256->3 + Linear 256->1 on every layer's features (regtr.py:399-443)."""

    def __init__(self, d_embed):
        super().__init__()
        self.coor_mlp = nn.Sequential(nn.Linear(d_embed, d_embed), nn.ReLU(), nn.Linear(d_embed, d_embed),
                                      nn.ReLU(), nn.Linear(d_embed, 3))
        self.conf_logits_decoder = nn.Linear(d_embed, 1)

    def forward_packed(self, feats):
        """feats (L,N,E) packed -> corr (L,N,3), logits (L,N,1)."""
        return self.coor_mlp(feats), self.conf_logits_decoder(feats)


class RegTR(nn.Module):
    def __init__(self, cfg, *args, **kwargs):
        super().__init__()
        self.cfg = cfg
        self.logger = logging.getLogger(self.__class__.__name__)
        self.preprocessor = PreprocessorGPU(cfg)
        self.kpf_encoder = KPFEncoder(cfg, cfg.d_embed)
        self.feat_proj = nn.Linear(self.kpf_encoder.encoder_skip_dims[-1], cfg.d_embed, bias=True)
        if cfg.get('pos_emb_type', 'sine') == 'sine':
            self.pos_embed = PositionEmbeddingCoordsSine(3, cfg.d_embed, scale=cfg.get('pos_emb_scaling', 1.0))
        else:
            raise NotImplementedError('learned position embedding is a "next" row (SURVEY.md 8f N4)')
        layer = TransformerCrossEncoderLayer(
            cfg.d_embed, cfg.nhead, cfg.d_feedforward, cfg.dropout, activation=cfg.transformer_act,
            normalize_before=cfg.pre_norm, sa_val_has_pos_emb=cfg.sa_val_has_pos_emb,
            ca_val_has_pos_emb=cfg.ca_val_has_pos_emb, attention_type=cfg.attention_type)
        norm = nn.LayerNorm(cfg.d_embed) if cfg.pre_norm else None
        self.transformer_encoder = TransformerCrossEncoder(layer, cfg.num_encoder_layers, norm,
                                                           return_intermediate=True)
        if cfg.get('direct_regress_coor', False):
            self.correspondence_decoder = CorrespondenceRegressor(cfg.d_embed)
        else:
            raise NotImplementedError('CorrespondenceDecoder (attention decoding) is a "next" row (SURVEY.md 8f N4)')
        if cfg.feature_loss_type == 'infonce':
            self.feature_criterion = _LossParam(cfg.d_embed)
            self.feature_criterion_un = _LossParam(cfg.d_embed)

    @property
    def device(self):
        return next(self.parameters()).device

    @torch.no_grad()
    def forward(self, batch):
        cfg = self.cfg
        B = len(batch['src_xyz'])
        # pyramid (regtr.py:117-122); a single host sync inside
        meta = self.preprocessor(list(batch['src_xyz']) + list(batch['tgt_xyz']))
        batch['kpconv_meta'] = meta
        lens_c = meta['_lens'][-1]
        feats0 = torch.ones_like(meta['points'][0][:, 0:1])
        # KPConv encoder -> bottleneck projection (regtr.py:136-146)
        feats_un, _ = self.kpf_encoder(feats0, meta)
        both_un = self.feat_proj(feats_un)
        # positions of the coarsest level and their embedding (regtr.py:149-154)
        xyz_c = meta['points'][-1]
        pe = self.pos_embed(xyz_c)
        # cross-encoder on packed tokens (regtr.py:156-166)
        plan = AttentionPlan(lens_c, xyz_c.device)
        cond = self.transformer_encoder.forward_packed(
            both_un.contiguous(), pe if cfg.transformer_encoder_has_pos_emb else None, plan)   # (L,N,E)
        # correspondence regression on all layers (regtr.py:168-171)
        corr, logit = self.correspondence_decoder.forward_packed(cond)                        # (L,N,3), (L,N,1)
        # correspondences + sigmoid + weighted Kabsch, one launch (regtr.py:185-203)
        pose = ops.pose_from_corr(xyz_c, corr.contiguous(), logit[..., 0].contiguous(), meta['_offs'][-1], B)

        split = lambda t, dim=0: torch.split(t, lens_c, dim=dim)
        un, kp = split(both_un), split(xyz_c)
        feat, warped, ovl = split(cond, 1), split(corr, 1), split(logit, 1)
        return {
            'src_feat_un': un[:B], 'tgt_feat_un': un[B:],
            'src_feat': list(feat[:B]), 'tgt_feat': list(feat[B:]),
            'src_kp': kp[:B], 'src_kp_warped': list(warped[:B]),
            'tgt_kp': kp[B:], 'tgt_kp_warped': list(warped[B:]),
            'src_overlap': list(ovl[:B]), 'tgt_overlap': list(ovl[B:]),
            'pose': pose,
        }
