"""CPU: host-side post-processing logic of the product (no kernels involved) against the oracle's restatement of
hipie_img.py:473-535 (panoptic merge) and :1025-1052 (token -> class pooling)."""
import types

import pytest
import torch


def _host_model():
    """A HIPIE_IMG shell with only the attributes the host-side helpers read (constructing the real one needs a GPU)."""
    from hipie_b200.modeling.hipie_img import HIPIE_IMG
    m = types.SimpleNamespace(object_mask_threshold=0.25, overlap_threshold=0.8)
    m.fused_sem_pano_finish = types.MethodType(HIPIE_IMG.fused_sem_pano_finish, m)
    m._pool_tables = types.MethodType(HIPIE_IMG._pool_tables, m)
    m.convert_grounding_to_od_logits = types.MethodType(HIPIE_IMG.convert_grounding_to_od_logits, m)
    return m


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_panoptic_merge_from_counters_matches_oracle(seed):
    """fused_sem_pano_finish consumes what the fused kernel emits (winner ids + per-query area counters); fed with those
    quantities computed by plain torch ops it must reproduce the oracle's panoptic_inference."""
    from hipie_oracle.model import HipieOracle
    g = torch.Generator().manual_seed(seed)
    Q, C, H, W = 40, 7, 48, 40
    logits = torch.randn(Q, H, W, generator=g) * 3 + torch.linspace(-2, 2, Q).view(Q, 1, 1)
    mask_cls = torch.softmax(torch.randn(Q, C, generator=g) * 4, -1)
    is_thing = {c + 1: c < 4 for c in range(C)}
    orc = types.SimpleNamespace(object_mask_threshold=0.25, overlap_threshold=0.8)
    ref_seg, ref_info = HipieOracle.panoptic_inference(orc, mask_cls, logits, is_thing)
    # what ops.seg_postprocess produces, in torch
    sg = logits.sigmoid()
    scores, labels = mask_cls.max(-1)
    keep = scores > 0.25
    prob = torch.where(keep.view(Q, 1, 1), scores.view(Q, 1, 1) * sg, torch.full_like(sg, -1.0))
    win = prob.argmax(0)
    any_keep = bool(keep.any())
    inter = sg.gather(0, win.unsqueeze(0))[0] >= 0.5
    ids = (2 * win + inter.long()).int() if any_keep else torch.full((H, W), -1, dtype=torch.int32)
    own = win.unsqueeze(0) == torch.arange(Q).view(Q, 1, 1)
    areas = torch.stack([own.flatten(1).sum(1), (sg >= 0.5).flatten(1).sum(1), (own & (sg >= 0.5)).flatten(1).sum(1)]).int()
    if not any_keep:
        areas.zero_()
    host = torch.cat([areas, labels.int().unsqueeze(0), keep.int().unsqueeze(0)])
    m = _host_model()
    sem, (seg, info) = m.fused_sem_pano_finish(dict(sem=None, ids=ids, Q=Q), host, is_thing)
    assert info == ref_info
    assert torch.equal(seg, ref_seg)


def test_token_to_class_tables_match_the_positive_map():
    """the pooling itself is a kernel now (hipie_class_scores, GPU test test_class_scores_matches_reference_loop); the host part is
    the table construction: token lists, counts (0 = class absent -> score 0 like the reference's zero-initialised tensor) and the
    FG / BG -9999 masks (hipie_img.py:1041-1049), cached by vocabulary CONTENT"""
    from hipie_oracle import synth
    _, _, pos_map, is_thing = synth.make_text(9, 64, seed=3)
    del pos_map[4]                                   # a class without tokens
    m = _host_model()
    tok, cnt, fg, bg = m._pool_tables(pos_map, 9, 64, is_thing, torch.device("cpu"))
    assert tok.dtype == torch.int32 and cnt.dtype == torch.int32 and fg.dtype == torch.int8
    for c in range(1, 10):
        if c in pos_map:
            assert tok[c - 1, :len(pos_map[c])].tolist() == pos_map[c] and int(cnt[c - 1]) == len(pos_map[c])
            assert bool(fg[c - 1]) == (not is_thing[c]) and bool(bg[c - 1]) == is_thing[c]
        else:
            assert int(cnt[c - 1]) == 0 and not bool(fg[c - 1]) and not bool(bg[c - 1])
    _, _, pos_map2, is_thing2 = synth.make_text(9, 64, seed=4)
    tok2, cnt2, _, _ = m._pool_tables(pos_map2, 9, 64, is_thing2, torch.device("cpu"))
    assert not (torch.equal(tok, tok2) and torch.equal(cnt, cnt2))
    assert m._pool_tables(pos_map, 9, 64, is_thing, torch.device("cpu"))[0] is tok      # cache hit on identical content
    with pytest.raises(ValueError):
        m._pool_tables({1: [70]}, 1, 64, {1: True}, torch.device("cpu"))                 # token index outside the text length
