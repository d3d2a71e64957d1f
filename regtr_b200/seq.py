"""Packed <-> padded sequence helpers.

The forward itself keeps tokens packed and needs none of these; they exist for callers written
against /root/reference/src/utils/seq_manipulation.py:6-48 (same names, arguments and return values), e.g.
code that feeds `TransformerCrossEncoder.forward` its padded (L,B,D) interface."""
from __future__ import annotations

import torch


def pad_sequence(sequences, require_padding_mask=False, require_lens=False, batch_first=False):
    """list of (Ni, D) -> padded (Nmax, B, D) [, bool mask (B, Nmax) True = padding] [, lengths]."""
    lens = [int(s.shape[0]) for s in sequences]
    n_max = max(lens) if lens else 0
    first = sequences[0]
    shape = (len(sequences), n_max) + tuple(first.shape[1:]) if batch_first else \
        (n_max, len(sequences)) + tuple(first.shape[1:])
    padded = first.new_zeros(shape)
    for b, s in enumerate(sequences):
        if batch_first:
            padded[b, :lens[b]] = s
        else:
            padded[:lens[b], b] = s
    mask = None
    if require_padding_mask:
        # indexed by padded.shape[0] like the reference (which is only meaningful for batch_first=False)
        mask = torch.arange(padded.shape[0], device=padded.device)[None, :] >= \
            torch.tensor(lens, device=padded.device)[:, None]
    return padded, mask, (lens if require_lens else None)


def unpad_sequences(padded, seq_lens):
    """([*,] Nmax, B, D) -> list of ([*,] Ni, D)."""
    return [padded[..., :seq_lens[b], b, :] for b in range(len(seq_lens))]


def split_src_tgt(feats, stack_lengths, dim=0):
    """Stacked (src_0..src_{B-1}, tgt_0..tgt_{B-1}) tensor -> (tuple of src parts, tuple of tgt parts)."""
    if isinstance(stack_lengths, torch.Tensor):
        stack_lengths = stack_lengths.tolist()
    B = len(stack_lengths) // 2
    parts = torch.split(feats, [int(v) for v in stack_lengths], dim=dim)
    return parts[:B], parts[B:]
