"""Oracle: decoupled MaskDINO branch — pixel decoder + DINO decoder + mask-embed contraction (test infra only).

Restates (H = /root/reference/projects/HIPIE/hipie/models/maskdino):
  H/pixel_decoder/maskdino_encoder.py:43-187,190-434     MSDeformAttnTransformerEncoderOnly, MaskDINOEncoder
  H/pixel_decoder/position_encoding.py:15-55             PositionEmbeddingSine (no -0.5 offset)
  H/transformer_decoder/maskdino_decoder.py:36-200,357-529  MaskDINODecoder (two-stage top-300, 9 layers, einsum)
  H/transformer_decoder/dino_decoder.py:18-270           TransformerDecoder, DeformableTransformerDecoderLayer
  H/utils/utils.py:11-100                                MLP, inverse_sigmoid, gen_encoder_output_proposals,
                                                         gen_sineembed_for_position
  H/meta_arch/maskdino_head.py:21-82                     MaskDINOHead
with the config of configs/mask_dino/maskdino_R50_bs16_50ep_3s_dowsample1_2048.yaml and the three input
features res3/res4/res5 HIPIE passes in (models/ddetrs_dn.py:172-186,863-888).
"""
import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .detr import MLP, FeatureResizer, MSDeformAttn, PositionEmbeddingSine, get_reference_points, inverse_sigmoid


def _get_clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


def gen_encoder_output_proposals(memory, memory_padding_mask, spatial_shapes):
    """utils/utils.py:33-71"""
    N_, S_, C_ = memory.shape
    proposals = []
    _cur = 0
    for lvl, (H_, W_) in enumerate(spatial_shapes):
        H_, W_ = int(H_), int(W_)
        mask_flatten_ = memory_padding_mask[:, _cur:(_cur + H_ * W_)].view(N_, H_, W_, 1)
        valid_H = torch.sum(~mask_flatten_[:, :, 0, 0], 1)
        valid_W = torch.sum(~mask_flatten_[:, 0, :, 0], 1)
        grid_y, grid_x = torch.meshgrid(torch.linspace(0, H_ - 1, H_, dtype=torch.float32),
                                        torch.linspace(0, W_ - 1, W_, dtype=torch.float32), indexing="ij")
        grid = torch.cat([grid_x.unsqueeze(-1), grid_y.unsqueeze(-1)], -1)
        scale = torch.cat([valid_W.unsqueeze(-1), valid_H.unsqueeze(-1)], 1).view(N_, 1, 1, 2)
        grid = (grid.unsqueeze(0).expand(N_, -1, -1, -1) + 0.5) / scale
        wh = torch.ones_like(grid) * 0.05 * (2.0 ** lvl)
        proposals.append(torch.cat((grid, wh), -1).view(N_, -1, 4))
        _cur += H_ * W_
    output_proposals = torch.cat(proposals, 1)
    output_proposals_valid = ((output_proposals > 0.01) & (output_proposals < 0.99)).all(-1, keepdim=True)
    output_proposals = torch.log(output_proposals / (1 - output_proposals))
    output_proposals = output_proposals.masked_fill(memory_padding_mask.unsqueeze(-1), float("inf"))
    output_proposals = output_proposals.masked_fill(~output_proposals_valid, float("inf"))
    output_memory = memory.masked_fill(memory_padding_mask.unsqueeze(-1), float(0))
    output_memory = output_memory.masked_fill(~output_proposals_valid, float(0))
    return output_memory, output_proposals


def gen_sineembed_for_position(pos_tensor):
    """utils/utils.py:74-100 -> [y, x, w, h]"""
    scale = 2 * math.pi
    dim_t = torch.arange(128, dtype=torch.float32)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="trunc") / 128)

    def emb(v):
        p = (v * scale)[:, :, None] / dim_t
        return torch.stack((p[:, :, 0::2].sin(), p[:, :, 1::2].cos()), dim=3).flatten(2)

    pos_x, pos_y = emb(pos_tensor[:, :, 0]), emb(pos_tensor[:, :, 1])
    if pos_tensor.size(-1) == 2:
        return torch.cat((pos_y, pos_x), dim=2)
    pos_w, pos_h = emb(pos_tensor[:, :, 2]), emb(pos_tensor[:, :, 3])
    return torch.cat((pos_y, pos_x, pos_w, pos_h), dim=2)


class MSDeformAttnTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model, d_ffn, n_levels, n_heads, n_points):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.norm2 = nn.LayerNorm(d_model)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        src2 = self.self_attn(src + pos, reference_points, src, spatial_shapes, level_start_index, padding_mask)
        src = self.norm1(src + src2)
        src2 = self.linear2(F.relu(self.linear1(src)))
        return self.norm2(src + src2)


class MSDeformAttnTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = _get_clones(encoder_layer, num_layers)

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None):
        output = src
        reference_points = get_reference_points(spatial_shapes, valid_ratios)
        for layer in self.layers:
            output = layer(output, pos, reference_points, spatial_shapes, level_start_index, padding_mask)
        return output


class MSDeformAttnTransformerEncoderOnly(nn.Module):
    """maskdino_encoder.py:43-115 (masks are always zeros for /32-divisible inputs, :82-88)."""

    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=2048, num_feature_levels=4, n_points=4):
        super().__init__()
        layer = MSDeformAttnTransformerEncoderLayer(d_model, dim_feedforward, num_feature_levels, nhead, n_points)
        self.encoder = MSDeformAttnTransformerEncoder(layer, num_encoder_layers)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        nn.init.normal_(self.level_embed)

    def forward(self, srcs, pos_embeds):
        masks = [torch.zeros((x.size(0), x.size(2), x.size(3)), dtype=torch.bool) for x in srcs]
        src_flatten, mask_flatten, lvl_pos_embed_flatten, spatial_shapes = [], [], [], []
        for lvl, (src, mask, pos_embed) in enumerate(zip(srcs, masks, pos_embeds)):
            bs, c, h, w = src.shape
            spatial_shapes.append((h, w))
            src_flatten.append(src.flatten(2).transpose(1, 2))
            mask_flatten.append(mask.flatten(1))
            lvl_pos_embed_flatten.append(pos_embed.flatten(2).transpose(1, 2) + self.level_embed[lvl].view(1, 1, -1))
        src_flatten = torch.cat(src_flatten, 1)
        mask_flatten = torch.cat(mask_flatten, 1)
        lvl_pos_embed_flatten = torch.cat(lvl_pos_embed_flatten, 1)
        spatial_shapes = torch.as_tensor(spatial_shapes, dtype=torch.long)
        level_start_index = torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
        valid_ratios = torch.ones(src_flatten.shape[0], len(srcs), 2)
        memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, lvl_pos_embed_flatten, mask_flatten)
        return memory, spatial_shapes, level_start_index


class ConvGN(nn.Conv2d):
    """detectron2 Conv2d(norm=GN(32), activation) wrapper (layers/wrappers.py:70-110): child `norm`."""

    def __init__(self, cin, cout, k, padding=0, relu=False):
        super().__init__(cin, cout, k, padding=padding, bias=False)
        self.norm = nn.GroupNorm(32, cout)
        self.relu = relu

    def forward(self, x):
        x = self.norm(F.conv2d(x, self.weight, None, self.stride, self.padding))
        return F.relu(x) if self.relu else x


class MaskDINOEncoder(nn.Module):
    """maskdino_encoder.py:190-434 with in_features res3/res4/res5 (channels given), feature_order low2high,
    conv_dim = mask_dim = 256, norm GN, common_stride 4 -> one extra FPN level on res3."""

    def __init__(self, in_channels, conv_dim=256, mask_dim=256, enc_layers=6, dim_ff=2048, nheads=8):
        super().__init__()
        c3, c4, c5 = in_channels
        proj = []
        for cin in (c3, c4, c5):                          # transformer_in_channels[::-1]  (:253-262)
            proj.append(nn.Sequential(nn.Conv2d(cin, conv_dim, kernel_size=1), nn.GroupNorm(32, conv_dim)))
        proj.append(nn.Sequential(nn.Conv2d(max(c3, c4, c5), conv_dim, kernel_size=3, stride=2, padding=1),
                                  nn.GroupNorm(32, conv_dim)))
        self.input_proj = nn.ModuleList(proj)
        for p in self.input_proj:
            nn.init.xavier_uniform_(p[0].weight, gain=1)
            nn.init.constant_(p[0].bias, 0)
        self.transformer = MSDeformAttnTransformerEncoderOnly(conv_dim, nheads, enc_layers, dim_ff, 4)
        self.pe_layer = PositionEmbeddingSine(conv_dim // 2, offset=0.0)
        self.mask_features = nn.Sequential(nn.ConvTranspose2d(conv_dim, conv_dim, 2, stride=2), nn.GroupNorm(32, conv_dim),
                                           nn.ReLU(), nn.Conv2d(conv_dim, mask_dim, kernel_size=1, stride=1, padding=0))
        self.adapter_1 = ConvGN(c3, conv_dim, 1)
        self.layer_1 = ConvGN(conv_dim, conv_dim, 3, padding=1, relu=True)

    def forward_features(self, features):
        f3, f4, f5 = features["res3"].float(), features["res4"].float(), features["res5"].float()
        zeros = lambda x: torch.zeros((x.size(0), x.size(2), x.size(3)), dtype=torch.bool)
        extra = self.input_proj[3](f5)
        srcs = [self.input_proj[0](f3), self.input_proj[1](f4), self.input_proj[2](f5), extra]
        pos = [self.pe_layer(f3, zeros(f3)), self.pe_layer(f4, zeros(f4)), self.pe_layer(f5, zeros(f5)),
               self.pe_layer(extra, zeros(extra))]
        y, spatial_shapes, level_start_index = self.transformer(srcs, pos)
        bs = y.shape[0]
        sizes = [int(h * w) for h, w in spatial_shapes]
        ys = torch.split(y, sizes, dim=1)
        out = [z.transpose(1, 2).view(bs, -1, int(spatial_shapes[i][0]), int(spatial_shapes[i][1])) for i, z in enumerate(ys)]
        cur_fpn = self.adapter_1(f3)
        yy = cur_fpn + F.interpolate(out[0], size=cur_fpn.shape[-2:], mode="bilinear", align_corners=False)
        out.append(self.layer_1(yy))
        return self.mask_features(out[-1]), out[0], out[:4]


class DinoDecoderLayer(nn.Module):
    """dino_decoder.py:171-270 (tensors are (nq, bs, d))."""

    def __init__(self, d_model, d_ffn, n_levels, n_heads, n_points):
        super().__init__()
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.norm1 = nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=0.0)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.norm3 = nn.LayerNorm(d_model)

    def forward(self, tgt, tgt_query_pos, tgt_reference_points, memory, memory_key_padding_mask, level_start_index,
                spatial_shapes):
        q = k = tgt + tgt_query_pos
        tgt2 = self.self_attn(q, k, tgt)[0]
        tgt = self.norm2(tgt + tgt2)
        tgt2 = self.cross_attn((tgt + tgt_query_pos).transpose(0, 1), tgt_reference_points.transpose(0, 1).contiguous(),
                               memory.transpose(0, 1), spatial_shapes, level_start_index, memory_key_padding_mask).transpose(0, 1)
        tgt = self.norm1(tgt + tgt2)
        tgt2 = self.linear2(F.relu(self.linear1(tgt)))
        return self.norm3(tgt + tgt2)


class DinoTransformerDecoder(nn.Module):
    """dino_decoder.py:18-168"""

    def __init__(self, decoder_layer, num_layers, norm, d_model=256):
        super().__init__()
        self.layers = _get_clones(decoder_layer, num_layers)
        self.norm = norm
        self.ref_point_head = MLP(2 * d_model, d_model, d_model, 2)
        self.bbox_embed = None
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()

    def forward(self, tgt, memory, memory_key_padding_mask, refpoints_unsigmoid, level_start_index, spatial_shapes,
                valid_ratios):
        output = tgt
        intermediate = []
        reference_points = refpoints_unsigmoid.sigmoid()
        ref_points = [reference_points]
        for layer_id, layer in enumerate(self.layers):
            reference_points_input = reference_points[:, :, None] * torch.cat([valid_ratios, valid_ratios], -1)[None, :]
            query_sine_embed = gen_sineembed_for_position(reference_points_input[:, :, 0, :])
            query_pos = self.ref_point_head(query_sine_embed)
            output = layer(output, query_pos, reference_points_input, memory, memory_key_padding_mask, level_start_index,
                           spatial_shapes)
            reference_before_sigmoid = inverse_sigmoid(reference_points)
            delta_unsig = self.bbox_embed[layer_id](output)
            new_reference_points = (delta_unsig + reference_before_sigmoid).sigmoid()
            reference_points = new_reference_points.detach()
            ref_points.append(new_reference_points)
            intermediate.append(self.norm(output))
        return [x.transpose(0, 1) for x in intermediate], [x.transpose(0, 1) for x in ref_points]


class MaskDINODecoder(nn.Module):
    """maskdino_decoder.py:36-200, forward :377-518, forward_prediction_heads :520-529 (eval, two_stage,
    initial_pred, dn inactive; class_embed is Linear(256 -> num_classes=256), ddetrs_dn.py:183-185)."""

    def __init__(self, hidden_dim=256, num_queries=300, nheads=8, dim_feedforward=2048, dec_layers=9, mask_dim=256,
                 num_classes=256, lang_dim=768):
        super().__init__()
        self.num_feature_levels = 4
        self.num_queries, self.num_layers, self.hidden_dim = num_queries, dec_layers, hidden_dim
        self.enc_output = nn.Linear(hidden_dim, hidden_dim)
        self.enc_output_norm = nn.LayerNorm(hidden_dim)
        self.class_embed = nn.Linear(hidden_dim, num_classes)
        self.resizer = FeatureResizer(lang_dim, hidden_dim)      # training-only (label enc), present in state_dict
        self.mask_embed = MLP(hidden_dim, hidden_dim, mask_dim, 3)
        self.decoder_norm = decoder_norm = nn.LayerNorm(hidden_dim)
        layer = DinoDecoderLayer(hidden_dim, dim_feedforward, 4, nheads, 4)
        self.decoder = DinoTransformerDecoder(layer, dec_layers, decoder_norm, hidden_dim)
        self._bbox_embed = _bbox_embed = MLP(hidden_dim, hidden_dim, 4, 3)
        nn.init.constant_(_bbox_embed.layers[-1].weight.data, 0)
        nn.init.constant_(_bbox_embed.layers[-1].bias.data, 0)
        self.bbox_embed = nn.ModuleList([_bbox_embed for _ in range(dec_layers)])
        self.decoder.bbox_embed = self.bbox_embed

    def forward_prediction_heads(self, output, mask_features, pred_mask=True):
        decoder_output = self.decoder_norm(output).transpose(0, 1)
        outputs_class = self.class_embed(decoder_output)
        outputs_mask = None
        if pred_mask:
            mask_embed = self.mask_embed(decoder_output)
            outputs_mask = torch.einsum("bqc,bchw->bqhw", mask_embed, mask_features)
        return outputs_class, outputs_mask

    def forward(self, x, mask_features, forced_topk=None):
        assert len(x) == self.num_feature_levels
        masks = [torch.zeros((s.size(0), s.size(2), s.size(3)), dtype=torch.bool) for s in x]
        src_flatten, mask_flatten, spatial_shapes = [], [], []
        for i in range(self.num_feature_levels):
            idx = self.num_feature_levels - 1 - i              # coarse -> fine (:398-404)
            spatial_shapes.append(x[idx].shape[-2:])
            src_flatten.append(x[idx].flatten(2).transpose(1, 2))
            mask_flatten.append(masks[i].flatten(1))
        src_flatten = torch.cat(src_flatten, 1)
        mask_flatten = torch.cat(mask_flatten, 1)
        spatial_shapes = torch.as_tensor(spatial_shapes, dtype=torch.long)
        level_start_index = torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
        valid_ratios = torch.ones(src_flatten.shape[0], 4, 2)
        output_memory, output_proposals = gen_encoder_output_proposals(src_flatten, mask_flatten, spatial_shapes)
        output_memory = self.enc_output_norm(self.enc_output(output_memory))
        enc_outputs_class_unselected = self.class_embed(output_memory)
        enc_outputs_coord_unselected = self._bbox_embed(output_memory) + output_proposals
        enc_scores = enc_outputs_class_unselected.max(-1)[0]
        topk_proposals = torch.topk(enc_scores, self.num_queries, dim=1)[1]
        if forced_topk is not None:
            topk_proposals = forced_topk
        refpoint_embed = torch.gather(enc_outputs_coord_unselected, 1, topk_proposals.unsqueeze(-1).repeat(1, 1, 4))
        tgt = torch.gather(output_memory, 1, topk_proposals.unsqueeze(-1).repeat(1, 1, self.hidden_dim))
        interm_class, interm_mask = self.forward_prediction_heads(tgt.transpose(0, 1), mask_features)   # dead at eval
        hs, references = self.decoder(tgt.transpose(0, 1), src_flatten.transpose(0, 1), mask_flatten,
                                      refpoint_embed.transpose(0, 1), level_start_index, spatial_shapes, valid_ratios)
        outputs_class, outputs_mask = self.forward_prediction_heads(hs[-1].transpose(0, 1), mask_features, True)
        # pred_box (:357-375): last layer box from reference[-2] and hs[-1]
        out_box = (self._bbox_embed(hs[-1]) + inverse_sigmoid(references[-2])).sigmoid()
        return {"pred_logits": outputs_class, "pred_masks": outputs_mask, "pred_boxes": out_box,
                "interm_masks": interm_mask, "enc_scores": enc_scores, "topk": topk_proposals}


class MaskDINOHead(nn.Module):
    def __init__(self, in_channels, hp):
        super().__init__()
        self.pixel_decoder = MaskDINOEncoder(in_channels, enc_layers=hp.get("md_enc_layers", 6), dim_ff=hp.get("md_dim_ff", 2048))
        self.predictor = MaskDINODecoder(num_queries=hp.get("md_queries", 300), dec_layers=hp.get("md_dec_layers", 9),
                                         dim_feedforward=hp.get("md_dim_ff", 2048), lang_dim=hp.get("lang_dim", 768))

    def forward(self, features, forced_topk=None):
        mask_features, _, multi_scale = self.pixel_decoder.forward_features(features)
        out = self.predictor(multi_scale, mask_features, forced_topk=forced_topk)
        out["mask_features"] = mask_features
        return out
