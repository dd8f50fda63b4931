"""CPU: host-side logic of the MaskCLIP path and of the BERT chunk planner (no kernels): prompt / offset / seen-class tables against the
oracle's restatement of hipie_img.py:818-830 and helper.py:112-130, the key-mask bit packing, the CLIP geometry read from a state_dict,
and the > 512-token cut plan against the oracle's BertEncoder (bert_model.py:68-135)."""
import types

import torch


def test_class_tables_match_oracle_helpers():
    from hipie_b200.modeling.maskclip import class_tables
    from hipie_oracle import clip as oc
    test_labels = [{"id": 1, "name": "person,child,girl"}, {"id": 2, "name": "wall"}, {"id": 3, "name": "zebra,okapi"}, {"id": 4, "name": "sky"}]
    train_labels = [{"id": 1, "name": "person,people"}, {"id": 2, "name": "sky,clouds"}]
    prompts, seg, overlap = class_tables(test_labels, train_labels, torch.device("cpu"))
    assert prompts == oc.prompt_labels([x["name"].split(",") for x in test_labels], "photo")
    assert seg.tolist() == [0, 3, 4, 6, 7] and seg.dtype == torch.int32
    assert overlap.tolist() == oc.category_overlapping_mask(test_labels, train_labels).tolist() == [1, 0, 0, 1]


def test_key_mask_bit_packing_roundtrip():
    from hipie_b200.modeling.maskclip import MaskCLIP
    g = torch.Generator().manual_seed(0)
    m = torch.rand(7, 96, generator=g) < 0.5
    m[0, 31] = True                                   # the sign bit of word 0
    m[1, 63] = True
    bits = MaskCLIP._pack_bits(m)
    assert bits.dtype == torch.int32 and bits.shape == (7, 3)
    b = bits.long() & 0xFFFFFFFF
    back = torch.stack([((b[:, k >> 5] >> (k & 31)) & 1).bool() for k in range(96)], 1)
    assert torch.equal(back, m)


def test_clip_geometry_from_state_dict():
    from hipie_b200.modeling.maskclip import config_from_state_dict
    from hipie_oracle import clip as oc
    for cfg in (oc.TINY,):
        assert config_from_state_dict(oc.CLIP(cfg).state_dict()) == cfg
    # the shipped default (MODEL.CLIP.NAME ViT-L-14-336): derive it from tensor SHAPES only (meta device, nothing allocated)
    with torch.device("meta"):
        big = oc.CLIP(oc.VIT_L_14_336)
    assert config_from_state_dict(big.state_dict()) == oc.VIT_L_14_336


def test_bert_chunk_plan_matches_oracle_chunking():
    """Engine.plan_text_chunks is pure host logic: its rows / masks / scatter spans must be what the oracle's BertEncoder builds
    for the same ids (the model forward itself is stubbed out on the oracle side to capture its inputs)."""
    from hipie_b200.modeling.engine import Engine
    from hipie_oracle import hparams, synth
    from hipie_oracle.model import BertEncoder
    hp = hparams.get("vit_tiny")
    _, ids, am = synth.make_batch(2, 64, 64, 230, 1024, seed=13)
    ids[1, 5:40] = ids[0, 300:335]                     # second row: a different prompt, different cut positions
    assert 512 < int(am[0].sum()) <= 1024
    fake = types.SimpleNamespace(device=torch.device("cpu"), hp=hp)
    plan = Engine.plan_text_chunks(fake, ids, am)
    enc = BertEncoder(hp)
    seen = {}

    def capture(input_ids=None, attention_mask=None, output_hidden_states=True):
        seen["ids"], seen["mask"] = input_ids.clone(), attention_mask.clone()
        h = torch.arange(input_ids.numel(), dtype=torch.float32).view(*input_ids.shape, 1).repeat(1, 1, 4)
        return types.SimpleNamespace(hidden_states=(None, h))
    del enc.model
    object.__setattr__(enc, "model", capture)
    out = enc({"input_ids": ids, "attention_mask": am})["hidden"]
    assert torch.equal(plan.rows, seen["ids"]) and torch.equal(plan.masks, seen["mask"].long())
    h = torch.arange(plan.rows.numel(), dtype=torch.float32).view(*plan.rows.shape, 1).repeat(1, 1, 4)
    mine = torch.zeros_like(out)
    for idx, (bi, s0, s1, t0, t1) in enumerate(plan.spans):
        mine[bi, t0:t1] = h[idx, s0:s1]
    assert torch.equal(mine, out)
    assert plan.signature[0:2] == (2, 1024) and len(plan.spans) >= 4
