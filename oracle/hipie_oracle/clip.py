"""CPU oracle of the MaskCLIP re-scoring branch -- TEST INFRASTRUCTURE, never imported by the product.

Two layers are restated here:

* the CLIP model the reference gets from its third-party dependency `open_clip_torch==2.0.2` (reference setup.py:198; the
  package is NOT vendored under /root/reference and is not installed in this image).  `QuickGELU`, `ResidualAttentionBlock`,
  `Transformer`, `VisualTransformer`, `CLIP` below restate the published architecture of open_clip 2.0.2 `model.py` (which is
  OpenAI CLIP's): pre-LN residual blocks over `nn.MultiheadAttention`, QuickGELU for `pretrained="openai"`, class token +
  learned positional embedding, `ln_pre` / `ln_post`, a bias-free projection; the text tower with the causal additive mask,
  `ln_final`, EOT pooling and `text_projection`.  Module / parameter names are open_clip's, so its state_dicts load.
  Parity for this layer is anchored on the reference's own call sites (open_vocab/clip.py reads `visual.conv1`,
  `.class_embedding`, `.positional_embedding`, `.ln_pre`, `.transformer`, `.ln_post`, `.proj`, `token_embedding`, `attn_mask`,
  `ln_final`, `text_projection`, `logit_scale`): tests/test_oracle_modules.py runs the UNMODIFIED reference `MaskCLIP` on top of
  this model through a stub `open_clip` module (oracle/ref_import.py) and holds `MaskCLIPOracle` to its outputs.
* the reference's own code: `MaskCLIP` (projects/HIPIE/hipie/open_vocab/clip.py:243-383), `ensemble_logits_with_labels` /
  `prompt_labels` (open_vocab/helper.py:79-130) and `HIPIE_IMG.get_clip_logits` (hipie_img.py:811-868).
"""
import copy
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)        # open_clip constants.py OPENAI_DATASET_MEAN / _STD
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)

# open_clip model_configs/ViT-L-14-336.json (the reference default MODEL.CLIP.NAME, config.py:156)
VIT_L_14_336 = dict(embed_dim=768, image_size=336, patch=14, width=1024, layers=24, heads=16, text_ctx=77, vocab=49408,
                    text_width=768, text_heads=12, text_layers=12)
# a small model with the same structure for tests (head width 64 like every CLIP ViT: clip.py:320 derives heads = width // 64)
TINY = dict(embed_dim=64, image_size=56, patch=14, width=128, layers=2, heads=2, text_ctx=16, vocab=512, text_width=64, text_heads=1,
            text_layers=2)


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head, mlp_ratio=4.0):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, int(d_model * mlp_ratio))), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(int(d_model * mlp_ratio), d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)

    def forward(self, x, attn_mask=None):
        y = self.ln_1(x)
        x = x + self.attn(y, y, y, need_weights=False, attn_mask=attn_mask)[0]
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads) for _ in range(layers)])

    def forward(self, x, attn_mask=None):
        for r in self.resblocks:
            x = r(x, attn_mask=attn_mask)
        return x


class VisualTransformer(nn.Module):
    def __init__(self, image_size, patch_size, width, layers, heads, output_dim):
        super().__init__()
        self.image_size = (image_size, image_size)
        self.grid_size = (image_size // patch_size, image_size // patch_size)
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(self.grid_size[0] * self.grid_size[1] + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))


class CLIP(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = dict(cfg)
        self.context_length = cfg["text_ctx"]
        self.visual = VisualTransformer(cfg["image_size"], cfg["patch"], cfg["width"], cfg["layers"], cfg["heads"], cfg["embed_dim"])
        self.transformer = Transformer(cfg["text_width"], cfg["text_layers"], cfg["text_heads"])
        self.vocab_size = cfg["vocab"]
        self.token_embedding = nn.Embedding(cfg["vocab"], cfg["text_width"])
        self.positional_embedding = nn.Parameter(0.01 * torch.randn(cfg["text_ctx"], cfg["text_width"]))
        self.ln_final = nn.LayerNorm(cfg["text_width"])
        self.text_projection = nn.Parameter(cfg["text_width"] ** -0.5 * torch.randn(cfg["text_width"], cfg["embed_dim"]))
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592600)            # log(1 / 0.07)
        mask = torch.empty(self.context_length, self.context_length).fill_(float("-inf")).triu_(1)
        self.register_buffer("attn_mask", mask, persistent=False)

    def encode_text(self, text):
        """open_clip CLIP.encode_text: EOT token = the highest id of each sequence."""
        x = self.token_embedding(text) + self.positional_embedding
        x = self.transformer(x.permute(1, 0, 2), attn_mask=self.attn_mask).permute(1, 0, 2)
        x = self.ln_final(x)
        return x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection


def ensemble_logits_with_labels(logits, labels, ensemble_method="max"):
    """helper.py:79-109: one logit per class = max (or mean) over the class's prompt synonyms."""
    len_list = [len(l) for l in labels]
    assert logits.shape[-1] == sum(len_list)
    out = torch.zeros(*logits.shape[:-1], len(labels), dtype=logits.dtype)
    for i in range(len(labels)):
        seg = logits[..., sum(len_list[:i]):sum(len_list[:i + 1])]
        out[..., i] = seg.max(dim=-1).values if ensemble_method == "max" else seg.mean(dim=-1)
    return out


def prompt_labels(labels, prompt="photo"):
    """helper.py:112-130."""
    labels = copy.deepcopy(labels)
    fmt = {"a": "a {}", "photo": "a photo of a {}.", "scene": "a photo of a {} in the scene."}[prompt]
    return [[fmt.format(l) for l in ls] for ls in labels]


class MaskCLIPOracle(nn.Module):
    """open_vocab/clip.py:243-383 on top of the CLIP model above (ClipAdapter :78-241 contributes `clip_preprocess`: the first
    two transforms, Resize / CenterCrop to the model's image size -- identities after get_mask_embed's own resize -- and the
    OpenAI normalisation)."""

    def __init__(self, clip):
        super().__init__()
        self.clip = clip.eval()
        self.register_buffer("mean", torch.tensor(OPENAI_MEAN).view(1, 3, 1, 1), persistent=False)
        self.register_buffer("std", torch.tensor(OPENAI_STD).view(1, 3, 1, 1), persistent=False)

    @property
    def logit_scale(self):                                      # :252-255
        return torch.clamp(self.clip.logit_scale.exp(), max=100)

    def attention_mask(self, mask):
        """:293-324: (B, Q, H, W) mask logits at the CLIP input size -> bool (B, Q + 1 + G, Q + 1 + G), True = masked out.
        Nobody attends to the mask tokens; mask token q attends to CLS and to the patches its mask touches."""
        v = self.clip.visual
        B, Q = mask.shape[:2]
        patch_mask = F.max_pool2d(mask.sigmoid(), kernel_size=v.conv1.kernel_size, stride=v.conv1.stride)
        tok = (patch_mask < 0.5).reshape(B, Q, -1)
        n_img = v.positional_embedding.shape[0]
        total = Q + n_img
        attn = torch.zeros((total, total), dtype=torch.bool)
        attn[:, :Q] = True
        attn = attn.unsqueeze(0).repeat_interleave(B, dim=0)
        attn[:, :Q, -(n_img - 1):] = tok
        return attn

    def mask_clip_forward(self, x, attn_mask, num_mask_tokens):
        """:257-286: mask tokens are copies of the (position-embedded, ln_pre'd) CLS token, prepended to the sequence."""
        v = self.clip.visual
        x = v.conv1(x)
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        x = torch.cat([v.class_embedding + torch.zeros(x.shape[0], 1, x.shape[-1]), x], dim=1)
        x = v.ln_pre(x + v.positional_embedding)
        x = x.permute(1, 0, 2)
        x = torch.cat([x[0:1].expand(num_mask_tokens, -1, -1), x], dim=0)
        x = v.transformer(x, attn_mask).permute(1, 0, 2)
        x = v.ln_post(x[:, :num_mask_tokens, :])
        return torch.einsum("nld,dc->nlc", x, v.proj)

    def get_mask_embed(self, image, mask):
        """:334-349 + :288-332.  image (B, 3, H, W) in 0..1; mask (B, Q, h, w) logits."""
        size = self.clip.visual.image_size
        image = F.interpolate(image, size=size, mode="bilinear", align_corners=False)
        mask = F.interpolate(mask, size=image.shape[-2:], mode="bilinear", align_corners=False)
        image = (image - self.mean) / self.std
        B, Q = mask.shape[:2]
        attn = self.attention_mask(mask)
        heads = self.clip.visual.conv1.out_channels // 64                     # :320 head width 64
        attn = attn.unsqueeze(1).expand(-1, heads, -1, -1).reshape(B * heads, attn.shape[-1], attn.shape[-1])
        return self.mask_clip_forward(image, attn, Q)

    def pred_logits(self, mask_embed, text_embed, labels):                   # :351-361
        logits = torch.einsum("bqc,nc->bqn", F.normalize(mask_embed, dim=-1), F.normalize(text_embed, dim=-1)) * self.logit_scale
        return ensemble_logits_with_labels(logits, labels)

    def build_text_embed(self, token_ids):
        """clip.py:29-73 after the tokenizer (open_clip.tokenize needs its BPE vocabulary file, which is not available offline:
        prompts arrive as token ids, (N, context_length) int64 with the EOT token as the highest id of each row)."""
        out = []
        for i in range(0, token_ids.shape[0], 256):
            out.append(self.clip.encode_text(token_ids[i:i + 256]))
        return torch.cat(out, 0)

    @torch.no_grad()
    def forward(self, image, mask, text_embed, labels):                      # :374-383
        mask_embed = self.get_mask_embed(image, mask)
        out = {"mask_embed": mask_embed}
        if text_embed is not None and labels is not None:
            out["mask_pred_open_logits"] = self.pred_logits(mask_embed, text_embed, labels)
        return out


def category_overlapping_mask(test_labels, train_labels):
    """hipie_img.py:818-830: 1 where a test class shares a synonym with any training class ("seen": weight alpha)."""
    train = {l for x in train_labels for l in x["name"].split(",")}
    return torch.tensor([int(not train.isdisjoint(set(x["name"].split(",")))) for x in test_labels], dtype=torch.long)


def fuse_clip_probs(pred_open_prob, mask_pred_open_logits, overlap, alpha=0.35, beta=0.7, agg_mode="MUL"):
    """hipie_img.py:840-866: geometric ('MUL') or arithmetic ('ADD') ensembling of the model's class probabilities with MaskCLIP's,
    weight alpha on seen classes and beta on unseen ones; returns log-probabilities ("pred_open_logits")."""
    if mask_pred_open_logits.shape[-1] == 1:
        p_clip = mask_pred_open_logits.sigmoid()
    else:
        p_clip = mask_pred_open_logits.softmax(dim=-1)
    if agg_mode == "ADD":
        base = (pred_open_prob * (1 - alpha) + p_clip * alpha + 1e-9).log() * overlap
        novel = (pred_open_prob * (1 - beta) + p_clip * beta + 1e-9).log() * (1 - overlap)
    else:
        base = (pred_open_prob ** (1 - alpha) * p_clip ** alpha).log() * overlap
        novel = (pred_open_prob ** (1 - beta) * p_clip ** beta).log() * (1 - overlap)
    return base + novel


def get_clip_logits(maskclip, test_labels, train_labels, mask_pred_results, image, pred_open_prob, text_embed, alpha=0.35, beta=0.7,
                    agg_mode="MUL"):
    """hipie_img.py:811-868 for one image.  `text_embed`: the cached prompt embeddings of `prompt_labels(test_labels, 'photo')`
    (one row per synonym, class-major), built by `build_text_embed`."""
    names = [x["name"].split(",") for x in test_labels]
    labels = prompt_labels(names, "photo")
    overlap = category_overlapping_mask(test_labels, train_labels)
    res = maskclip(image, mask_pred_results, text_embed, labels)
    return fuse_clip_probs(pred_open_prob, res["mask_pred_open_logits"][0], overlap, alpha, beta, agg_mode)


def init_clip_(clip, seed=0):
    """Deterministic CLIP weights keyed by parameter name (synth.fill_by_name_), with the two non-random pieces restored: the causal
    additive mask buffer and logit_scale = log(1 / 0.07) (open_clip's initial value; OpenAI's trained one is ~4.6 -> exp clamps at 100)."""
    from .synth import fill_by_name_
    fill_by_name_(clip, seed)
    with torch.no_grad():
        clip.attn_mask.copy_(torch.empty_like(clip.attn_mask).fill_(float("-inf")).triu_(1))
        clip.logit_scale.fill_(2.6592600)
    return clip.eval()


def synth_clip_tokenize(texts, context_length, vocab_size):
    """Stand-in for open_clip.tokenize (its BPE vocabulary file is not available offline): [SOT] + one id per whitespace word (a
    CRC of the word) + [EOT], zero padded; SOT = vocab - 2 and EOT = vocab - 1, so EOT is the highest id of the row as in CLIP
    (encode_text pools at argmax).  Deterministic across processes."""
    import zlib
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, t in enumerate(texts):
        ids = [vocab_size - 2] + [zlib.crc32(w.encode()) % (vocab_size - 3) + 1 for w in t.lower().split()][:context_length - 2] + [vocab_size - 1]
        out[i, :len(ids)] = torch.tensor(ids)
    return out
