"""CPU: pin the oracle restatement against fixtures produced by the REAL reference leaf modules
(oracle/gen_golden.py imported them from /root/reference in the build container)."""
import os

import torch

from hipie_oracle import detr, maskdino, vit


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name))


def test_vit_utils(golden_dir):
    g = _load(golden_dir, "vit_utils.pt")
    win, pad_hw = vit.window_partition(g["x"], 14)
    assert torch.equal(win, g["win"]) and tuple(pad_hw) == tuple(g["pad_hw"])
    assert torch.equal(vit.window_unpartition(win, 14, pad_hw, (20, 24)), g["back"])
    assert torch.equal(vit.get_rel_pos(6, 6, g["rph"]), g["Rh"])          # exercises the linear interpolation branch
    assert torch.equal(vit.get_rel_pos(5, 5, g["rpw"]), g["Rw"])
    out = vit.add_decomposed_rel_pos(g["attn"].clone(), g["q"], g["rph"], g["rpw"], (6, 5), (6, 5))
    assert torch.allclose(out, g["attn_out"], atol=1e-6)
    assert torch.allclose(vit.get_abs_pos(g["abs_pos"], True, (20, 24)), g["abs_pos_out"], atol=1e-6)


def test_position_encodings_and_maskdino_utils(golden_dir):
    g = _load(golden_dir, "posenc_utils.pt")
    x = torch.zeros(2, 4, 9, 11)
    pos_detr = detr.PositionEmbeddingSine(128, offset=-0.5)(x, g["mask"])
    assert torch.allclose(pos_detr, g["pos_detr"], atol=1e-6)
    pos_md = detr.PositionEmbeddingSine(128, offset=0.0)(x, torch.zeros(2, 9, 11, dtype=torch.bool))
    assert torch.allclose(pos_md, g["pos_md"], atol=1e-6)
    om, op = maskdino.gen_encoder_output_proposals(g["mem"], torch.zeros(2, g["mem"].shape[1], dtype=torch.bool), g["ss"])
    assert torch.equal(om, g["out_mem"]) and torch.equal(op, g["out_prop"])
    assert torch.allclose(maskdino.gen_sineembed_for_position(g["pos4"]), g["sine4"], atol=1e-6)
    assert torch.equal(detr.inverse_sigmoid(g["inv_sig_in"]), g["inv_sig_out"])
    # the DETR-side sine embedding of 4-d reference points equals the MaskDINO one ([y, x, w, h] order)
    assert torch.allclose(detr.get_sine_pos_embed(g["pos4"]), g["sine4"], atol=1e-5)


def test_vl_fusion_block(golden_dir):
    g = _load(golden_dir, "vlfuse_block.pt")
    blk = detr.BiAttentionBlockForCheckpoint(32, 48, 64, 4, init_values=1.0 / 6).eval()
    missing = blk.load_state_dict(g["state"], strict=True)
    with torch.no_grad():
        ov, ol = blk(g["v"], g["l"], attention_mask_l=g["mask"])
    assert torch.allclose(ov, g["out_v"], atol=1e-6) and torch.allclose(ol, g["out_l"], atol=1e-6)
