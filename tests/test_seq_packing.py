"""Coarse-level packing vs golden vectors generated from the reference's own Python
(tests/golden/make_golden_seq.py): the list-based mirrors split_src_tgt / pad_sequence / unpad_sequences (CPU), and
-m gpu: the kernel-backed pad_stacked, PositionEmbeddingCoordsSine and pack_coarse_level (= the whole coarse-level step
of RegTR.forward, finegrained_regtr.py:149-172).  Indexing is exact; the embedding holds 1e-6; the projection 1e-4."""
import os

import numpy as np
import pytest
import torch

import kpreg_b200  # noqa: F401
from kpreg_b200.position_embedding import PositionEmbeddingCoordsSine, PositionEmbeddingLearned
from kpreg_b200.seq_manipulation import pad_sequence, pad_stacked, split_src_tgt, unpad_sequences
from conftest import ROOT


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "seq_packing.npz")))


def test_split_pad_unpad_match_reference(g):
    feats, lens = torch.from_numpy(g["feats"]), torch.from_numpy(g["lens"])
    src, tgt = split_src_tgt(feats, lens)
    assert [int(s.shape[0]) for s in src] == g["lens"][:3].tolist() and [int(t.shape[0]) for t in tgt] == g["lens"][3:].tolist()
    src_pad, src_mask, src_lens = pad_sequence(src, require_padding_mask=True, require_lens=True)
    tgt_pad, tgt_mask, none_lens = pad_sequence(tgt, require_padding_mask=True)
    assert none_lens is None and src_lens == g["lens"][:3].tolist()
    assert np.array_equal(src_pad.numpy(), g["src_pad"]) and np.array_equal(src_mask.numpy(), g["src_mask"])
    assert np.array_equal(tgt_pad.numpy(), g["tgt_pad"]) and np.array_equal(tgt_mask.numpy(), g["tgt_mask"])
    assert src_mask.dtype == torch.bool
    assert np.array_equal(unpad_sequences(src_pad, src_lens)[1].numpy(), g["unpad_1"])
    padded, mask, lens_out = pad_sequence(src)           # defaults: no mask, no lengths (reference :20-21)
    assert mask is None and lens_out is None and padded.shape == src_pad.shape
    bf, bf_mask, _ = pad_sequence(src, require_padding_mask=True, batch_first=True)
    assert np.array_equal(bf.numpy(), g["src_pad"].transpose(1, 0, 2)) and np.array_equal(bf_mask.numpy(), g["src_mask"])


@pytest.mark.gpu
def test_pad_stacked_is_split_then_pad(g):
    feats, lens = torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["lens"]).cuda()
    for max_len in (None, (int(g["lens"][:3].max()), int(g["lens"][3:].max()))):
        s, t, sm, tm = pad_stacked(feats, lens, max_len)
        assert np.array_equal(s.cpu().numpy(), g["src_pad"]) and np.array_equal(t.cpu().numpy(), g["tgt_pad"])
        assert np.array_equal(sm.cpu().numpy(), g["src_mask"]) and np.array_equal(tm.cpu().numpy(), g["tgt_mask"])
        assert sm.dtype == torch.bool
    # ragged extremes: an empty cloud, a single point
    lens2 = torch.tensor([0, 3, 1, 2])
    f2 = torch.arange(12, dtype=torch.float32).reshape(6, 2)
    s, t, sm, tm = pad_stacked(f2.cuda(), lens2.cuda())
    want_s, want_m, _ = pad_sequence(split_src_tgt(f2, lens2)[0], require_padding_mask=True)
    want_t, want_tm, _ = pad_sequence(split_src_tgt(f2, lens2)[1], require_padding_mask=True)
    assert torch.equal(s.cpu(), want_s) and torch.equal(sm.cpu(), want_m) and torch.equal(t.cpu(), want_t) and torch.equal(tm.cpu(), want_tm)
    with pytest.raises(RuntimeError):
        pad_stacked(f2, lens2)  # CPU tensors: no CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("d_model,scale", [(256, 1.0), (64, 0.5), (100, 1.0)])
def test_sine_position_embedding_matches_reference(g, d_model, scale):
    emb = PositionEmbeddingCoordsSine(3, d_model, scale=scale)(torch.from_numpy(g["xyz"]).cuda()).cpu()
    want = g[f"sine_{d_model}"]
    assert emb.shape == want.shape
    assert float(np.abs(emb.numpy() - want).max()) < 1e-6
    pad = d_model - (d_model // 3 // 2 * 2) * 3
    if pad:
        assert float(emb[:, -pad:].abs().max()) == 0.0  # unused dimensions are zero
    batched = PositionEmbeddingCoordsSine(3, d_model, scale=scale)(torch.from_numpy(g["xyz"]).cuda()[None, :59])  # (*, d_in) -> (*, d_out)
    assert batched.shape == (1, 59, d_model) and torch.equal(batched[0].cpu(), emb[:59])


@pytest.mark.gpu
def test_pack_coarse_level_matches_the_reference_sequence(g):
    """feat_proj -> split_src_tgt -> pad_sequence(+mask) and pos_embed -> split -> pad, as RegTR.forward spells them
    (finegrained_regtr.py:149-172), against the outputs of exactly that reference code."""
    from kpreg_b200.seq_manipulation import pack_coarse_level
    proj = torch.nn.Linear(16, 24, bias=True)
    proj.load_state_dict({"weight": torch.from_numpy(g["proj_w"]), "bias": torch.from_numpy(g["proj_b"])})
    proj = proj.cuda()
    pos_embed = PositionEmbeddingCoordsSine(3, 24, scale=1.0)
    feats, xyz, lens = (torch.from_numpy(g[k]).cuda() for k in ("feats", "xyz", "lens"))
    with torch.no_grad():
        for max_len in (None, (52, 41)):
            out = pack_coarse_level(feats, xyz, lens, proj, pos_embed, max_len)
            for half in ("src", "tgt"):
                f, want_f = out[f"{half}_feats_padded"].cpu().numpy(), g[f"regtr_{half}_feats_padded"]
                assert f.shape == want_f.shape and float(np.abs(f - want_f).max()) < 1e-4 * float(np.abs(want_f).max())
                assert np.array_equal(f == 0, want_f == 0)              # padded rows are exactly zero
                pe, want_pe = out[f"{half}_pe_padded"].cpu().numpy(), g[f"regtr_{half}_pe_padded"]
                assert pe.shape == want_pe.shape and float(np.abs(pe - want_pe).max()) < 1e-6
                assert np.array_equal(out[f"{half}_key_padding_mask"].cpu().numpy(), g[f"regtr_{half}_mask"])
        # a learned position embedding goes through the generic route (evaluate, then pad like the features)
        learned = PositionEmbeddingLearned(3, 24).cuda()
        out = pack_coarse_level(feats, xyz, lens, proj, learned, (52, 41))
        want = torch.nn.utils.rnn.pad_sequence(list(torch.split(learned.mlp(xyz), g["lens"].tolist())[:3]))
        assert float((out["src_pe_padded"] - want).abs().max()) < 1e-4 * float(want.abs().max())


def test_learned_position_embedding_state_dict_is_the_reference_layout():
    sd = PositionEmbeddingLearned(3, 256).state_dict()
    assert list(sd) == [f"mlp.{i}.{p}" for i in (0, 2, 4, 6, 8) for p in ("weight", "bias")]
    assert [tuple(sd[f"mlp.{i}.weight"].shape) for i in (0, 2, 4, 6, 8)] == [(32, 3), (64, 32), (128, 64), (256, 128), (256, 256)]
