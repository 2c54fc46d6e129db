"""Coarse-level sequence packing between the encoder and the transformer — host-side mirror of the reference's
``utils/seq_manipulation.py`` (:6-48), same names and return conventions, used at ``models/finegrained_regtr.py:149-186``.

These are pure indexing operations on device tensors (the step immediately after the hot path).
``pad_stacked`` / ``pack_coarse_level`` are the batched forms of what the reference spells as split -> pad per list
(plus the projection and the position embedding): one CUDA kernel over the whole stacked level, no per-cloud Python
loop and no list of views.  They need CUDA tensors; the list-based mirrors below run wherever their inputs live.
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


def _max_lens(lens: torch.Tensor, max_len):
    b = lens.shape[0] // 2
    if max_len is not None:
        return int(max_len[0]), int(max_len[1])
    if b == 0:
        return 0, 0
    m = torch.stack([lens[:b].max(), lens[b:].max()]).tolist()   # the one read-back that sizing the outputs needs
    return int(m[0]), int(m[1])


def pad_stacked(feats: torch.Tensor, stack_lengths: torch.Tensor, max_len: Optional[Sequence[int]] = None
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Stacked (N, D) features of 2B clouds (sources first, then targets) -> (src_padded (Ns_max, B, D), tgt_padded
    (Nt_max, B, D), src_mask (B, Ns_max), tgt_mask (B, Nt_max)) with masks True at padded positions: the result of
    ``pad_sequence(split_src_tgt(feats, lens)[k], require_padding_mask=True)`` for both halves, from ONE kernel
    (kpreg_pack_coarse) instead of a split, two pads and a Python loop over clouds for the masks.
    ``max_len`` = (Ns_max, Nt_max) avoids the one device read-back that sizing the outputs needs."""
    from . import ops
    lens = stack_lengths.to(device=feats.device)
    ns_max, nt_max = _max_lens(lens, max_len)
    src, tgt, _, _, sm, tm = ops.pack_coarse(feats, None, lens, (ns_max, nt_max), int(feats.shape[1]))
    return src, tgt, sm, tm


def pack_coarse_level(feats_un: torch.Tensor, xyz: torch.Tensor, stack_lengths: torch.Tensor, feat_proj, pos_embed,
                      max_len: Optional[Sequence[int]] = None):
    """The step between the KPConv encoder and the transformer (reference models/finegrained_regtr.py:149-172):

        both = feat_proj(feats_un);  src, tgt = split_src_tgt(both, lens);  pe = split_src_tgt(pos_embed(xyz), lens)
        *_pe_padded = pad_sequence(pe);  *_feats_padded, *_key_padding_mask = pad_sequence(feats, require_padding_mask=True)

    as two launches: the projection (tcgen05 Linear with the bias in its epilogue) and kpreg_pack_coarse, which writes the
    padded features, computes the sine position code of the coarse points straight into its padded layout and emits
    both padding masks.  ``feat_proj``: nn.Linear; ``pos_embed``: PositionEmbeddingCoordsSine (any other module is
    evaluated on the stacked points and padded like the features).  Returns a dict with the reference's variable names."""
    from . import ops
    from .position_embedding import PositionEmbeddingCoordsSine
    lens = stack_lengths.to(device=feats_un.device)
    ns_max, nt_max = _max_lens(lens, max_len)
    both = ops.linear_forward(feats_un, feat_proj.weight, None, feat_proj.bias)
    d = int(both.shape[1])
    if isinstance(pos_embed, PositionEmbeddingCoordsSine) and pos_embed.n_dim == 3 and pos_embed.d_model == d:
        sf, tf, sp, tp, sm, tm = ops.pack_coarse(both, xyz, lens, (ns_max, nt_max), d, pos_embed.num_pos_feats, pos_embed.scale,
                                                 pos_embed.frequencies(xyz.device))
    else:
        sf, tf, _, _, sm, tm = ops.pack_coarse(both, None, lens, (ns_max, nt_max), d)
        pe = pos_embed(xyz)
        sp, tp, _, _, _, _ = ops.pack_coarse(pe, None, lens, (ns_max, nt_max), int(pe.shape[1]))
    return {'src_feats_padded': sf, 'tgt_feats_padded': tf, 'src_key_padding_mask': sm, 'tgt_key_padding_mask': tm,
            'src_pe_padded': sp, 'tgt_pe_padded': tp, 'both_feats_un': both}
