"""Coarse-level sequence packing between the encoder and the transformer — host-side mirror of the reference's
``utils/seq_manipulation.py`` (:6-48), same names and return conventions, used at ``models/finegrained_regtr.py:149-186``.

These are pure indexing operations on device tensors (the step immediately after the hot path).
``pad_stacked`` is the batched form the reference spells as split -> pad per list: one gather for the whole stacked
level, no per-cloud Python loop and no list of views.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def split_src_tgt(feats: torch.Tensor, stack_lengths, dim: int = 0):
    """(sources, targets): the first B and the last B chunks of ``feats`` split by ``stack_lengths`` (reference :42-48)."""
    if isinstance(stack_lengths, torch.Tensor):
        stack_lengths = stack_lengths.tolist()
    b = len(stack_lengths) // 2
    separate = torch.split(feats, list(stack_lengths), dim=dim)
    return separate[:b], separate[b:]


def pad_sequence(sequences: Sequence[torch.Tensor], require_padding_mask: bool = False, require_lens: bool = False,
                 batch_first: bool = False):
    """List of (N_i, D) sequences -> (padded (N_max, B, D) [or (B, N_max, D)], padding_mask (B, N_max) bool with True at
    padded positions or None, lengths or None) — reference :6-33.  The mask is one comparison, not a loop over clouds."""
    padded = torch.nn.utils.rnn.pad_sequence(list(sequences), batch_first=batch_first)
    padding_mask, padding_lens = None, None
    if require_padding_mask:
        n_max = padded.shape[1] if batch_first else padded.shape[0]
        lens = torch.tensor([int(s.shape[0]) for s in sequences], device=padded.device)
        padding_mask = torch.arange(n_max, device=padded.device)[None, :] >= lens[:, None]
    if require_lens:
        padding_lens = [int(s.shape[0]) for s in sequences]
    return padded, padding_mask, padding_lens


def unpad_sequences(padded: torch.Tensor, seq_lens: Sequence[int]) -> List[torch.Tensor]:
    """Reverse of pad_sequence for (..., N_max, B, D) tensors (reference :36-39)."""
    return [padded[..., :seq_lens[b], b, :] for b in range(len(seq_lens))]


def pad_stacked(feats: torch.Tensor, stack_lengths: torch.Tensor, max_len: Optional[int] = None
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Stacked (N, D) features of 2B clouds (sources first, then targets) -> (src_padded (Ns_max, B, D), tgt_padded
    (Nt_max, B, D), src_mask (B, Ns_max), tgt_mask (B, Nt_max)) with masks True at padded positions: the result of
    ``pad_sequence(split_src_tgt(feats, lens)[k], require_padding_mask=True)`` for both halves, from two gathers.
    ``max_len`` = (Ns_max, Nt_max) avoids the one device read-back that sizing the outputs needs."""
    lens = stack_lengths.to(device=feats.device, dtype=torch.long)
    b = lens.shape[0] // 2
    starts = torch.cumsum(lens, 0) - lens
    if max_len is None:
        m = torch.stack([lens[:b].max(), lens[b:].max()]).tolist() if b > 0 else [0, 0]
    else:
        m = list(max_len)
    out = []
    for half, n_max in ((slice(0, b), int(m[0])), (slice(b, 2 * b), int(m[1]))):
        pos = torch.arange(n_max, device=feats.device)[:, None]                       # (N_max, 1)
        valid = pos < lens[half][None, :]                                              # (N_max, B)
        rows = (starts[half][None, :] + pos).clamp(max=max(feats.shape[0] - 1, 0))     # (N_max, B)
        padded = feats[rows] * valid[..., None].to(feats.dtype) if feats.shape[0] > 0 else feats.new_zeros((n_max, b, feats.shape[1]))
        out.append((padded, ~valid.t()))
    return out[0][0], out[1][0], out[0][1], out[1][1]
