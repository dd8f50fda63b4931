"""CPU: the oracle's module-level restatements against outputs of the REAL reference modules
(oracle/gen_golden_modules.py ran the unmodified files of /root/reference in the build container; see its docstring and
tests/golden/MANIFEST.md for what each fixture pins).  Weights are re-created by name with hipie_oracle.synth.fill_by_name_,
so every test also checks that the oracle module exposes exactly the reference's parameter names and shapes."""
import copy
import os

import pytest
import torch
import torch.nn as nn

from hipie_oracle import detr, maskdino, vit
from hipie_oracle.synth import fill_by_name_


def _load(golden_dir, name):
    p = os.path.join(golden_dir, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not generated")
    return torch.load(p)


def _same_keys(module, keys):
    mine = sorted(module.state_dict().keys())
    assert mine == sorted(keys), (sorted(set(mine) ^ set(keys))[:10])


def test_vit_forward_matches_reference(golden_dir):
    """backbone/vit.py ViT.forward: patch embed + bicubic abs-pos + windowed/global blocks with decomposed rel-pos + FPN"""
    g = _load(golden_dir, "ref_vit.pt")
    m = vit.ViT(**g["kw"]).eval()
    _same_keys(m, g["keys"])
    fill_by_name_(m, g["seed"])
    with torch.no_grad():
        out = m(g["x"])
    for k in ("res3", "res4", "res5"):
        ref = g["out"][k]          # fp32 rounding only: 1e-5 of the feature range
        assert (out[k] - ref).abs().max() < 1e-5 * ref.abs().max() + 1e-5, (k, (out[k] - ref).abs().max(), ref.abs().max())


def _oracle_transformer(kw):
    tr = detr.DeformableTransformerVLDINO(**kw)
    d, nd = kw["d_model"], kw["num_decoder_layers"]
    class_embed = detr.VL_Align(kw["lang_dim"], d)
    bbox_embed = detr.MLP(d, d, 4, 3)
    tr.decoder.class_embed = nn.ModuleList([copy.deepcopy(class_embed) for _ in range(nd + 1)])
    tr.decoder.bbox_embed = nn.ModuleList([copy.deepcopy(bbox_embed) for _ in range(nd + 1)])
    tr.decoder.class_embed[-1] = detr.Still_Classifier(d)
    return tr


def test_transformer_forward_matches_reference(golden_dir):
    """deformable_transformer_dino.py DeformableTransformerVLDINO.forward with padded inputs: VL fusion, 2 encoder layers
    (MSDeformAttn.forward), two-stage proposals + top-k, bg queries, 2 decoder layers, intermediate reference points"""
    g = _load(golden_dir, "ref_transformer.pt")
    tr = _oracle_transformer(g["kw"]).eval()
    _same_keys(tr, g["keys"])
    fill_by_name_(tr, g["seed"])
    lang = {k: v.clone() for k, v in g["lang"].items()}
    with torch.no_grad():
        hs, memory, init_ref, inter_refs, lang_out, aux = tr(g["srcs"], g["masks"], g["poses"], lang)
    assert torch.allclose(memory, g["memory"], atol=2e-5), (memory - g["memory"]).abs().max()
    assert torch.allclose(lang_out["hidden"], g["lang_hidden"], atol=2e-5)
    assert torch.allclose(aux["enc_scores"], g["enc_cls"][..., 0], atol=2e-5)
    assert torch.allclose(init_ref, g["init_ref"], atol=1e-5)
    assert torch.allclose(hs, g["hs"], atol=5e-5), (hs - g["hs"]).abs().max()
    assert torch.allclose(inter_refs, g["inter_refs"], atol=1e-5)


def test_heads_match_reference(golden_dir):
    """deformable_detr.py VL_Align.forward (normalise, /2, projection, per-token bias, exp(log_scale), clamp) and MLP"""
    g = _load(golden_dir, "ref_heads.pt")
    va = fill_by_name_(detr.VL_Align(768, 256), g["va_seed"]).eval()
    with torch.no_grad():
        va.log_scale.fill_(0.3)
        assert torch.allclose(va(g["q"], g["emb"]), g["va_out"], atol=1e-5)
        mlp = fill_by_name_(detr.MLP(256, 256, 4, 3), g["mlp_seed"]).eval()
        assert torch.allclose(mlp(g["q"]), g["mlp_out"], atol=1e-5)
