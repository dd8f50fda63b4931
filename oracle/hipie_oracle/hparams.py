"""Hyper-parameter sets for the oracle (test infrastructure).  The full-size values come from
projects/HIPIE/configs/training/{r50,vit_huge_32g}.yaml, hipie/config.py and the MaskDINO yaml."""
import copy

BERT_BASE = dict(vocab=30522, hidden=768, layers=12, heads=12, inter=3072, max_pos=512)

VIT_H = dict(backbone="vit",
             vit=dict(embed_dim=1280, depth=32, num_heads=16, window_size=14, window_block_indexes=(0, 1, 3, 4, 6, 7, 9, 10),
                      img_size=1024, patch_size=16, pretrain_img_size=224),
             hidden_dim=256, enc_layers=6, dec_layers=6, dim_ff=2048, num_queries=900, num_bg=10, vl_hidden=2048, lang_dim=768,
             md_queries=300, md_dec_layers=9, md_enc_layers=6, md_dim_ff=2048, bert=BERT_BASE, max_query_len=512)

R50 = dict(VIT_H, backbone="r50")
R50.pop("vit")

# tiny configurations for parity tests: same code paths (windowed + global ViT blocks with rel-pos interpolation,
# 4 feature levels, two-stage top-k, bg queries, MaskDINO branch), seconds on CPU.
VIT_TINY = dict(backbone="vit",
                vit=dict(embed_dim=160, depth=4, num_heads=2, window_size=14, window_block_indexes=(0, 1, 3),
                         img_size=256, patch_size=16, pretrain_img_size=224),
                hidden_dim=256, enc_layers=2, dec_layers=2, dim_ff=512, num_queries=60, num_bg=4, vl_hidden=2048, lang_dim=768,
                md_queries=30, md_dec_layers=2, md_enc_layers=2, md_dim_ff=512,
                bert=dict(vocab=30522, hidden=768, layers=2, heads=12, inter=3072, max_pos=512), max_query_len=64)


def get(name):
    return copy.deepcopy({"vit_h": VIT_H, "r50": R50, "vit_tiny": VIT_TINY}[name])
