"""Golden fixtures from the REAL reference MODULES (not just leaf helpers), run in the build container:

    python oracle/gen_golden_modules.py            # writes tests/golden/ref_*.pt

oracle/ref_import.py mounts /root/reference/projects/HIPIE/hipie so that the unmodified reference files execute here
(their absent third-party packages are stubbed; the MSDeformAttn pybind op is bound to the reference's own
ms_deform_attn_core_pytorch).  Each fixture = seeded inputs + the reference module's state_dict + its outputs at a small
size; tests/test_oracle_modules.py loads the state_dict into the oracle's restatement of the same module and requires the
same outputs, which pins these rows of SURVEY.md §8a to the reference itself:

  ref_vit.pt          a2-a4   backbone/vit.py ViT.forward (PatchEmbed, abs-pos, windowed + global Block/Attention, FPN)
  ref_transformer.pt  a8-a13  deformable_transformer_dino.py DeformableTransformerVLDINO.forward (VLFuse, encoder layers with
                              MSDeformAttn.forward, two-stage proposals, bg queries, decoder layers, look-forward-twice refs)
  ref_heads.pt        a14-a15 deformable_detr.py VL_Align / Still_Classifier / MLP
  ref_maskdino.pt     a17-a19 maskdino_encoder.py MaskDINOEncoder.forward_features, maskdino_decoder.py MaskDINODecoder.forward
                              (+ dino_decoder.py TransformerDecoder, forward_prediction_heads with the mask-embed einsum)
  ref_condinst.pt     a16     ddetrs_dn.py dynamic_mask_with_coords / mask_heads_forward / parse_dynamic_params /
                              compute_locations / aligned_bilinear, MaskHeadSmallConv.forward
  ref_bert_chunk.pt   a7      bert_model.py BertEncoder.forward incl. the > 512-token chunk path (random-init HF BertModel)
  ref_postproc.pt     a20-a23 hipie_img.py convert_grounding_to_od_logits / semantic_inference / panoptic_inference,
                              ddetrs.py segmentation_postprocess
"""
import copy
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
import ref_import  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


from hipie_oracle.synth import fill_by_name_  # noqa: E402


def randomize_(module, seed):
    """weights = a deterministic function of (parameter name, shape, seed): the fixtures carry inputs and outputs only and the
    test re-creates the same weights in the oracle's restatement through the same names (hipie_oracle.synth.fill_by_name_)"""
    return fill_by_name_(module, seed)


def small_cfg():
    from hipie_b200.config import add_hipie_config, get_cfg
    cfg = get_cfg()
    add_hipie_config(cfg)
    m = cfg.MODEL
    m.DECOUPLE_TGT, m.STILL_TGT_FOR_BOTH, m.VL_FUSION_USE_CHECKPOINT = True, True, False
    m.DDETRS.HIDDEN_DIM, m.DDETRS.VL_HIDDEN_DIM, m.DDETRS.ENC_LAYERS, m.DDETRS.DEC_LAYERS = 256, 64, 2, 2
    m.DDETRS.TWO_STAGE_NUM_BG_PROPOSALS, m.DDETRS.NUM_VL_LAYERS = 3, 1
    m.LANGUAGE_BACKBONE.LANG_DIM = 768      # the transformer's resizer hard-codes 768 (deformable_transformer_dino.py:99-103)
    return cfg


def gen_vit():
    vit = ref_import.ref("backbone.vit")
    torch.manual_seed(1)
    kw = dict(img_size=64, patch_size=16, embed_dim=32, depth=3, num_heads=2, use_rel_pos=True, window_size=14,
              window_block_indexes=(0, 2), pretrain_img_size=224, out_feature="last_feat")
    m = randomize_(vit.ViT(**kw), 2).eval()
    x = torch.randn(2, 3, 96, 80)
    with torch.no_grad():
        out = m(x)
    torch.save(dict(kw=dict(embed_dim=32, depth=3, num_heads=2, window_size=14, window_block_indexes=(0, 2), img_size=64, patch_size=16,
                            pretrain_img_size=224),
                    seed=2, keys=sorted(m.state_dict().keys()), x=x, out=out), os.path.join(OUT, "ref_vit.pt"))
    print("ref_vit.pt", {k: tuple(v.shape) for k, v in out.items()})


def _attach_heads(tr, dd, cfg, d, nd):
    """what DeformableDETR.__init__ does to the transformer (deformable_detr.py:262-292): per-layer VL_Align / MLP clones,
    Still_Classifier as the encoder-proposal scorer"""
    class_embed = dd.VL_Align(cfg)
    bbox_embed = dd.MLP(d, d, 4, 3)
    tr.decoder.class_embed = nn.ModuleList([copy.deepcopy(class_embed) for _ in range(nd + 1)])
    tr.decoder.bbox_embed = nn.ModuleList([copy.deepcopy(bbox_embed) for _ in range(nd + 1)])
    tr.decoder.class_embed[-1] = dd.Still_Classifier(d)


def gen_transformer():
    cfg = small_cfg()
    tmod = ref_import.ref("models.deformable_detr.deformable_transformer_dino")
    dd = ref_import.ref("models.deformable_detr.deformable_detr")
    torch.manual_seed(3)
    d, nd = 256, 2      # 256: the decoder's ref_point_head consumes 4 x 128 sine features (hard-coded)
    tr = tmod.DeformableTransformerVLDINO(d_model=d, nhead=8, num_encoder_layers=2, num_decoder_layers=nd, dim_feedforward=40, dropout=0.0,
                                          activation="relu", return_intermediate_dec=True, num_feature_levels=4, dec_n_points=4,
                                          enc_n_points=4, two_stage=True, two_stage_num_proposals=12, look_forward_twice=True,
                                          mixed_selection=True, use_checkpoint=False, cfg=cfg)
    _attach_heads(tr, dd, cfg, d, nd)
    randomize_(tr, 4).eval()
    g = torch.Generator().manual_seed(5)
    shapes = [(12, 10), (6, 5), (3, 3), (2, 2)]
    srcs = [torch.randn(2, d, h, w, generator=g) for h, w in shapes]
    masks = []
    for h, w in shapes:
        mk = torch.zeros(2, h, w, dtype=torch.bool)
        mk[1, h - max(1, h // 4):, :] = True          # image 1 is padded at the bottom and on the right
        mk[1, :, w - max(1, w // 5):] = True
        masks.append(mk)
    masks[3][1] = False
    masks[3][1, 1:, :] = True
    poses = [torch.randn(2, d, h, w, generator=g) for h, w in shapes]
    lang = {"hidden": torch.randn(2, 9, 768, generator=g), "masks": torch.ones(2, 9, dtype=torch.long)}
    lang["masks"][1, 6:] = 0
    lang_in = {k: v.clone() for k, v in lang.items()}
    with torch.no_grad():
        hs, memory, init_ref, inter_refs, enc_cls, enc_coord, lang_out = tr(
            srcs, masks, poses, query_embed=(None, None), mask_on=True, language_dict_features=lang, task="detection")
    torch.save(dict(seed=4, keys=sorted(tr.state_dict().keys()), srcs=srcs, masks=masks, poses=poses, lang=lang_in, hs=hs, memory=memory, init_ref=init_ref,
                    inter_refs=inter_refs, enc_cls=enc_cls, enc_coord=enc_coord, lang_hidden=lang_out["hidden"],
                    kw=dict(d_model=d, nhead=8, num_encoder_layers=2, num_decoder_layers=nd, dim_feedforward=40, two_stage_num_proposals=12,
                            num_bg=3, vl_hidden=64, lang_dim=768)),
               os.path.join(OUT, "ref_transformer.pt"))
    print("ref_transformer.pt hs", tuple(hs.shape), "memory", tuple(memory.shape))
    # heads
    torch.manual_seed(6)
    va = randomize_(dd.VL_Align(cfg), 7).eval()
    with torch.no_grad():
        va.log_scale.fill_(0.3)
    q = torch.randn(2, 5, d, generator=g)
    emb = torch.randn(2, 9, 768, generator=g)
    mlp = randomize_(dd.MLP(d, d, 4, 3), 8).eval()
    with torch.no_grad():
        torch.save(dict(va_seed=7, mlp_seed=8, q=q, emb=emb, va_out=va(q, emb), mlp_out=mlp(q)),
                   os.path.join(OUT, "ref_heads.pt"))
    print("ref_heads.pt")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["vit", "transformer", "maskdino", "condinst", "bert", "postproc"]
    for w in which:
        globals()["gen_" + w]()
