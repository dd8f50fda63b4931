"""Oracle: Deformable-DETR (DINO variant) with VL early fusion (test infrastructure only).

Restates, for the eval path with the flags every shipped yaml sets (SURVEY.md §8a notes):
  H = /root/reference/projects/HIPIE/hipie
  H/models/deformable_detr/position_encoding.py:20-56          PositionEmbeddingSine (-0.5 variant)
  H/models/deformable_detr/ops/modules/ms_deform_attn.py:30-116 MSDeformAttn
  H/models/deformable_detr/fuse_helper.py:7-179, vlfusion.py:64-120  BiMultiHeadAttention / VLFuse
  H/models/deformable_detr/deformable_transformer_dino.py:28-670     transformer, encoder, decoder, MLP, sine embed
  H/models/deformable_detr/deformable_detr.py:40-82,193-292          VL_Align, Still_Classifier, DeformableDETRDINO
  H/util/misc.py:493-497 inverse_sigmoid
"""
import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .msda import ms_deform_attn_core


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    x1 = x.clamp(min=eps)
    x2 = (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)


def _get_clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = F.relu(layer(x)) if i < self.num_layers - 1 else layer(x)
        return x


class FeatureResizer(nn.Module):
    """deformable_transformer_dino.py:578-597 (eps 1e-12); dropout is identity at eval."""

    def __init__(self, input_feat_size, output_feat_size):
        super().__init__()
        self.fc = nn.Linear(input_feat_size, output_feat_size, bias=True)
        self.layer_norm = nn.LayerNorm(output_feat_size, eps=1e-12)

    def forward(self, x):
        return self.layer_norm(self.fc(x))


class PositionEmbeddingSine(nn.Module):
    """position_encoding.py:20-56, normalize=True, num_pos_feats=128, (cumsum - 0.5) variant."""

    def __init__(self, num_pos_feats=128, temperature=10000, offset=-0.5):
        super().__init__()
        self.num_pos_feats, self.temperature, self.scale, self.offset = num_pos_feats, temperature, 2 * math.pi, offset

    def forward(self, x, mask):
        not_mask = ~mask
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        eps = 1e-6
        y_embed = (y_embed + self.offset) / (y_embed[:, -1:, :] + eps) * self.scale
        x_embed = (x_embed + self.offset) / (x_embed[:, :, -1:] + eps) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32)
        dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / self.num_pos_feats)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
        pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
        return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


class MSDeformAttn(nn.Module):
    """ms_deform_attn.py:30-116 with the core op = ms_deform_attn_core_pytorch (the MaskDINO copy's own
    CPU fallback, maskdino/pixel_decoder/ops/modules/ms_deform_attn.py:116-121)."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        nn.init.constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2).repeat(
            1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid_init[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid_init.view(-1))
        nn.init.constant_(self.attention_weights.weight.data, 0.)
        nn.init.constant_(self.attention_weights.bias.data, 0.)
        nn.init.xavier_uniform_(self.value_proj.weight.data)
        nn.init.constant_(self.value_proj.bias.data, 0.)
        nn.init.xavier_uniform_(self.output_proj.weight.data)
        nn.init.constant_(self.output_proj.bias.data, 0.)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in
        value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        sampling_offsets = self.sampling_offsets(query).view(N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        attention_weights = self.attention_weights(query).view(N, Len_q, self.n_heads, self.n_levels * self.n_points)
        attention_weights = F.softmax(attention_weights, -1).view(N, Len_q, self.n_heads, self.n_levels, self.n_points)
        if reference_points.shape[-1] == 2:
            offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            sampling_locations = reference_points[:, :, None, :, None, :] \
                + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            sampling_locations = reference_points[:, :, None, :, None, :2] \
                + sampling_offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError("Last dim of reference_points must be 2 or 4")
        output = ms_deform_attn_core(value, input_spatial_shapes, sampling_locations, attention_weights)
        return self.output_proj(output)


class BiMultiHeadAttention(nn.Module):
    """fuse_helper.py:7-139 (eval: dropout off; clamps on, STABLE_SOFTMAX_2D off)."""

    def __init__(self, v_dim, l_dim, embed_dim, num_heads):
        super().__init__()
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.scale = self.head_dim ** (-0.5)
        self.v_proj = nn.Linear(v_dim, embed_dim)
        self.l_proj = nn.Linear(l_dim, embed_dim)
        self.values_v_proj = nn.Linear(v_dim, embed_dim)
        self.values_l_proj = nn.Linear(l_dim, embed_dim)
        self.out_v_proj = nn.Linear(embed_dim, v_dim)
        self.out_l_proj = nn.Linear(embed_dim, l_dim)
        self._reset_parameters()

    def _reset_parameters(self):
        for m in (self.v_proj, self.l_proj, self.values_v_proj, self.values_l_proj, self.out_v_proj, self.out_l_proj):
            nn.init.xavier_uniform_(m.weight)
            m.bias.data.fill_(0)

    def _shape(self, tensor, seq_len, bsz):
        return tensor.view(bsz, seq_len, self.num_heads, self.head_dim).transpose(1, 2).contiguous()

    def forward(self, v, l, attention_mask_l=None):
        bsz, tgt_len, _ = v.size()
        query_states = self.v_proj(v) * self.scale
        key_states = self._shape(self.l_proj(l), -1, bsz)
        value_v_states = self._shape(self.values_v_proj(v), -1, bsz)
        value_l_states = self._shape(self.values_l_proj(l), -1, bsz)
        proj_shape = (bsz * self.num_heads, -1, self.head_dim)
        query_states = self._shape(query_states, tgt_len, bsz).view(*proj_shape)
        key_states = key_states.view(*proj_shape)
        value_v_states = value_v_states.view(*proj_shape)
        value_l_states = value_l_states.view(*proj_shape)
        src_len = key_states.size(1)
        attn_weights = torch.bmm(query_states, key_states.transpose(1, 2))
        attn_weights = torch.clamp(attn_weights, min=-50000)
        attn_weights = torch.clamp(attn_weights, max=50000)
        attn_weights_T = attn_weights.transpose(1, 2)
        attn_weights_l = attn_weights_T - torch.max(attn_weights_T, dim=-1, keepdim=True)[0]
        attn_weights_l = torch.clamp(attn_weights_l, min=-50000)
        attn_weights_l = torch.clamp(attn_weights_l, max=50000)
        attn_weights_l = attn_weights_l.softmax(dim=-1)
        if attention_mask_l is not None:
            assert attention_mask_l.dim() == 2
            attention_mask = attention_mask_l.unsqueeze(1).unsqueeze(1)
            attention_mask = attention_mask.expand(bsz, 1, tgt_len, src_len)
            attention_mask = attention_mask.masked_fill(attention_mask == 0, -9e15)   # valid tokens keep "+1"
            attn_weights = attn_weights.view(bsz, self.num_heads, tgt_len, src_len) + attention_mask
            attn_weights = attn_weights.view(bsz * self.num_heads, tgt_len, src_len)
        attn_weights_v = F.softmax(attn_weights, dim=-1)
        attn_output_v = torch.bmm(attn_weights_v, value_l_states)
        attn_output_l = torch.bmm(attn_weights_l, value_v_states)
        attn_output_v = attn_output_v.view(bsz, self.num_heads, tgt_len, self.head_dim).transpose(1, 2).reshape(
            bsz, tgt_len, self.embed_dim)
        attn_output_l = attn_output_l.view(bsz, self.num_heads, src_len, self.head_dim).transpose(1, 2).reshape(
            bsz, src_len, self.embed_dim)
        return self.out_v_proj(attn_output_v), self.out_l_proj(attn_output_l)


class BiAttentionBlockForCheckpoint(nn.Module):
    """fuse_helper.py:142-179: residual is taken on the *normalised* inputs."""

    def __init__(self, v_dim, l_dim, embed_dim, num_heads, init_values):
        super().__init__()
        self.layer_norm_v = nn.LayerNorm(v_dim)
        self.layer_norm_l = nn.LayerNorm(l_dim)
        self.attn = BiMultiHeadAttention(v_dim, l_dim, embed_dim, num_heads)
        self.gamma_v = nn.Parameter(init_values * torch.ones(v_dim))
        self.gamma_l = nn.Parameter(init_values * torch.ones(l_dim))

    def forward(self, v, l, attention_mask_l=None):
        v = self.layer_norm_v(v)
        l = self.layer_norm_l(l)
        delta_v, delta_l = self.attn(v, l, attention_mask_l=attention_mask_l)
        return v + self.gamma_v * delta_v, l + self.gamma_l * delta_l


class VLFuse(nn.Module):
    """vlfusion.py:64-120; mutates the language dict in place (:114-115)."""

    def __init__(self, img_dim, lang_dim, embed_dim, enc_layers):
        super().__init__()
        self.b_attn = BiAttentionBlockForCheckpoint(img_dim, lang_dim, embed_dim, 8, init_values=1.0 / enc_layers)

    def forward(self, x):
        lang = x["lang"]
        fused_v, fused_l = self.b_attn(x["visual"], lang["hidden"], lang["masks"])
        lang["hidden"] = fused_l
        return {"visual": fused_v, "lang": lang}


class DeformableTransformerEncoderLayer(nn.Module):
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


def get_reference_points(spatial_shapes, valid_ratios):
    """deformable_transformer_dino.py:313-325"""
    reference_points_list = []
    for lvl, (H_, W_) in enumerate(spatial_shapes):
        H_, W_ = int(H_), int(W_)
        ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32),
                                      torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32), indexing="ij")
        ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
        ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
        reference_points_list.append(torch.stack((ref_x, ref_y), -1))
    reference_points = torch.cat(reference_points_list, 1)
    return reference_points[:, :, None] * valid_ratios[:, None]


class DeformableTransformerEncoderVL(nn.Module):
    """:302-351 — VLFuse only on the first NUM_VL_LAYERS (=1) layers; lang_layers are Identity."""

    def __init__(self, vl_fusion_layer, encoder_layer, num_layers, num_vl_layers):
        super().__init__()
        self.vl_layers = nn.ModuleList([copy.deepcopy(vl_fusion_layer) if i < num_vl_layers else nn.Identity()
                                        for i in range(num_layers)])
        self.layers = _get_clones(encoder_layer, num_layers)
        self.lang_layers = nn.ModuleList([nn.Identity() for _ in range(num_layers)])

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos, padding_mask, language_dict_features):
        output = {"visual": src, "lang": language_dict_features}
        reference_points = get_reference_points(spatial_shapes, valid_ratios)
        for vl_layer, layer in zip(self.vl_layers, self.layers):
            output = vl_layer(output)
            output["visual"] = layer(output["visual"], pos, reference_points, spatial_shapes, level_start_index, padding_mask)
        return output


class DeformableTransformerDecoderLayer(nn.Module):
    def __init__(self, d_model, d_ffn, n_levels, n_heads, n_points):
        super().__init__()
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.norm1 = nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=0.0)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.norm3 = nn.LayerNorm(d_model)

    def forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, level_start_index, src_padding_mask=None):
        q = k = tgt + query_pos
        tgt2 = self.self_attn(q.transpose(0, 1), k.transpose(0, 1), tgt.transpose(0, 1))[0].transpose(0, 1)
        tgt = self.norm2(tgt + tgt2)
        tgt2 = self.cross_attn(tgt + query_pos, reference_points, src, src_spatial_shapes, level_start_index, src_padding_mask)
        tgt = self.norm1(tgt + tgt2)
        tgt2 = self.linear2(F.relu(self.linear1(tgt)))
        return self.norm3(tgt + tgt2)


def get_sine_pos_embed(pos_tensor, num_pos_feats=128, temperature=10000, exchange_xy=True):
    """:636-670 -> [y, x, w, h] order"""
    scale = 2 * math.pi
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)

    def sine_func(x):
        sin_x = x * scale / dim_t
        return torch.stack((sin_x[:, :, 0::2].sin(), sin_x[:, :, 1::2].cos()), dim=3).flatten(2)

    pos_res = [sine_func(x) for x in pos_tensor.split([1] * pos_tensor.shape[-1], dim=-1)]
    if exchange_xy:
        pos_res[0], pos_res[1] = pos_res[1], pos_res[0]
    return torch.cat(pos_res, dim=2)


class DeformableTransformerDecoder(nn.Module):
    """:453-525 (return_intermediate, look_forward_twice, box refine)."""

    def __init__(self, embed_dim, decoder_layer, num_layers):
        super().__init__()
        self.layers = _get_clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.ref_point_head = MLP(2 * embed_dim, embed_dim, embed_dim, 2)
        self.bbox_embed = None
        self.class_embed = None

    def forward(self, tgt, reference_points, src, src_spatial_shapes, src_level_start_index, src_valid_ratios,
                src_padding_mask=None):
        output = tgt
        intermediate, intermediate_reference_points = [], []
        for lid, layer in enumerate(self.layers):
            reference_points_input = reference_points[:, :, None] * torch.cat([src_valid_ratios, src_valid_ratios], -1)[:, None]
            query_sine_embed = get_sine_pos_embed(reference_points_input[:, :, 0, :])
            query_pos = self.ref_point_head(query_sine_embed)
            output = layer(output, query_pos, reference_points_input, src, src_spatial_shapes, src_level_start_index,
                           src_padding_mask)
            tmp = self.bbox_embed[lid](output)
            new_reference_points = (tmp + inverse_sigmoid(reference_points)).sigmoid()
            reference_points = new_reference_points.detach()
            intermediate.append(output)
            intermediate_reference_points.append(new_reference_points)
        return torch.stack(intermediate), torch.stack(intermediate_reference_points)


def agg_lang_feat(features, mask):
    """:28-43 (average)"""
    embedded = features * mask.unsqueeze(-1).float()
    return embedded.sum(1) / (mask.sum(-1).unsqueeze(-1).float())


class DeformableTransformerVLDINO(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048,
                 num_feature_levels=4, n_points=4, two_stage_num_proposals=900, num_bg=10, vl_hidden=2048, lang_dim=768):
        super().__init__()
        self.d_model, self.nhead = d_model, nhead
        self.two_stage_num_proposals = two_stage_num_proposals
        enc_layer = DeformableTransformerEncoderLayer(d_model, dim_feedforward, num_feature_levels, nhead, n_points)
        vl = VLFuse(d_model, lang_dim, vl_hidden, num_encoder_layers)
        self.encoder = DeformableTransformerEncoderVL(vl, enc_layer, num_encoder_layers, num_vl_layers=1)
        dec_layer = DeformableTransformerDecoderLayer(d_model, dim_feedforward, num_feature_levels, nhead, n_points)
        self.decoder = DeformableTransformerDecoder(d_model, dec_layer, num_decoder_layers)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self.tgt_embed = nn.Embedding(two_stage_num_proposals, d_model)
        self.tgt_embed_bg = nn.Embedding(num_bg, d_model)
        self.bg_query_refs = nn.Embedding(num_bg, 4)
        self.enc_output = nn.Linear(d_model, d_model)
        self.enc_output_norm = nn.LayerNorm(d_model)
        self._reset_parameters()
        self.resizer = FeatureResizer(lang_dim, d_model)   # dead compute at eval (multiplied by 0.0, :259-261)

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, (MSDeformAttn, BiMultiHeadAttention)):
                m._reset_parameters()
        nn.init.normal_(self.level_embed)

    def gen_encoder_output_proposals(self, memory, memory_padding_mask, spatial_shapes):
        """:138-168"""
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
        output_memory = self.enc_output_norm(self.enc_output(output_memory))
        return output_memory, output_proposals

    @staticmethod
    def get_valid_ratio(mask):
        _, H, W = mask.shape
        valid_H = torch.sum(~mask[:, :, 0], 1)
        valid_W = torch.sum(~mask[:, 0, :], 1)
        return torch.stack([valid_W.float() / W, valid_H.float() / H], -1)

    def forward(self, srcs, masks, pos_embeds, language_dict_features, forced_topk=None):
        """:180-299 (mask_on=True, query_embed=(None,None), DECOUPLE_TGT & STILL_TGT_FOR_BOTH).
        forced_topk: optional (B, K) indices that replace the discontinuous top-k (parity harness)."""
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
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
        vl = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, lvl_pos_embed_flatten, mask_flatten,
                          language_dict_features)
        memory, language_dict_features = vl["visual"], vl["lang"]
        bs = memory.shape[0]
        output_memory, output_proposals = self.gen_encoder_output_proposals(memory, mask_flatten, spatial_shapes)
        enc_outputs_class = self.decoder.class_embed[self.decoder.num_layers](output_memory)
        enc_outputs_coord_unact = self.decoder.bbox_embed[self.decoder.num_layers](output_memory) + output_proposals
        topk_proposals = torch.topk(enc_outputs_class[..., 0], self.two_stage_num_proposals, dim=1)[1]
        if forced_topk is not None:
            topk_proposals = forced_topk
        topk_coords_unact = torch.gather(enc_outputs_coord_unact, 1, topk_proposals.unsqueeze(-1).repeat(1, 1, 4))
        reference_points = topk_coords_unact.sigmoid()
        tgt = self.tgt_embed.weight[None].repeat(bs, 1, 1)
        bg_query = self.tgt_embed_bg.weight[None].repeat(bs, 1, 1)
        tgt = torch.cat([bg_query, tgt], dim=1)
        bg_ref_pt = self.bg_query_refs.weight[None].repeat(bs, 1, 1)
        reference_points = torch.cat([bg_ref_pt, reference_points], dim=1)
        init_reference_out = reference_points
        hs, inter_references = self.decoder(tgt, reference_points, memory, spatial_shapes, level_start_index, valid_ratios,
                                            src_padding_mask=mask_flatten)
        aux = dict(enc_scores=enc_outputs_class[..., 0], topk=topk_proposals, spatial_shapes=spatial_shapes,
                   level_start_index=level_start_index, valid_ratios=valid_ratios, mask_flatten=mask_flatten)
        return hs, memory, init_reference_out, inter_references, language_dict_features, aux


class VL_Align(nn.Module):
    """deformable_detr.py:40-73 (PRIOR_PROB 0.01, LOG_SCALE 0.0, CLAMP_DOT_PRODUCT on)."""

    def __init__(self, lang_dim=768, hidden_dim=256, prior_prob=0.01, log_scale=0.0):
        super().__init__()
        bias_value = -math.log((1 - prior_prob) / prior_prob)
        self.dot_product_projection_text = nn.Linear(lang_dim, hidden_dim, bias=True)
        self.log_scale = nn.Parameter(torch.Tensor([log_scale]))
        self.bias_lang = nn.Parameter(torch.zeros(lang_dim))
        self.bias0 = nn.Parameter(torch.Tensor([bias_value]))

    def forward(self, x, embedding):
        embedding = F.normalize(embedding, p=2, dim=-1)
        dot_product_proj_tokens = self.dot_product_projection_text(embedding / 2.0)
        dot_product_proj_tokens_bias = torch.matmul(embedding, self.bias_lang) + self.bias0
        A = x.shape[1]
        bias = dot_product_proj_tokens_bias.unsqueeze(1).repeat(1, A, 1)
        logit = (torch.matmul(x, dot_product_proj_tokens.transpose(-1, -2)) / self.log_scale.exp()) + bias
        logit = torch.clamp(logit, max=50000)
        return torch.clamp(logit, min=-50000)


class Still_Classifier(nn.Module):
    def __init__(self, hidden_dim):
        super().__init__()
        self.body = nn.Linear(hidden_dim, 1)

    def forward(self, x, lang_feat=None):
        return self.body(x)


class Joiner(nn.Sequential):
    """backbone.py:112-129 with MaskedBackbone (masked_backbone.py:10-43): index 0 wraps the backbone as
    `.backbone`, index 1 is the position embedding."""

    def __init__(self, backbone, position_embedding):
        wrapper = nn.Module()
        wrapper.backbone = backbone
        super().__init__(wrapper, position_embedding)
        self.num_channels = backbone.num_channels
        self.strides = backbone.strides

    def forward(self, tensors, mask):
        xs = self[0].backbone(tensors)
        out, pos = [], []
        for name in sorted(xs.keys()):
            x = xs[name]
            m = F.interpolate(mask[None].float(), size=x.shape[-2:]).to(torch.bool)[0]
            out.append((x, m))
        for x, m in out:
            pos.append(self[1](x, m).to(x.dtype))
        return out, pos


class DeformableDETRDINO(nn.Module):
    """deformable_detr.py:193-292 (two_stage, with_box_refine, USE_IOU_BRANCH, STILL_CLS_FOR_ENCODER)."""

    def __init__(self, backbone, transformer, num_feature_levels=4, lang_dim=768):
        super().__init__()
        self.transformer = transformer
        hidden_dim = transformer.d_model
        class_embed = VL_Align(lang_dim, hidden_dim)
        bbox_embed = MLP(hidden_dim, hidden_dim, 4, 3)
        iou_head = nn.Linear(hidden_dim, 1)
        self.num_feature_levels = num_feature_levels
        num_backbone_outs = len(backbone.strides)
        input_proj_list = []
        for i in range(num_backbone_outs):
            in_channels = backbone.num_channels[i]
            input_proj_list.append(nn.Sequential(nn.Conv2d(in_channels, hidden_dim, kernel_size=1), nn.GroupNorm(32, hidden_dim)))
        for _ in range(num_feature_levels - num_backbone_outs):
            input_proj_list.append(nn.Sequential(nn.Conv2d(in_channels, hidden_dim, kernel_size=3, stride=2, padding=1),
                                                 nn.GroupNorm(32, hidden_dim)))
            in_channels = hidden_dim
        self.input_proj = nn.ModuleList(input_proj_list)
        self.backbone = backbone
        bias_value = -math.log((1 - 0.01) / 0.01)
        iou_head.bias.data = torch.ones(1) * bias_value
        nn.init.constant_(bbox_embed.layers[-1].weight.data, 0)
        nn.init.constant_(bbox_embed.layers[-1].bias.data, 0)
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)
        num_pred = transformer.decoder.num_layers + 1
        self.class_embed = _get_clones(class_embed, num_pred)
        self.bbox_embed = _get_clones(bbox_embed, num_pred)
        self.iou_head = _get_clones(iou_head, num_pred - 1)
        nn.init.constant_(self.bbox_embed[0].layers[-1].bias.data[2:], -2.0)
        self.transformer.decoder.bbox_embed = self.bbox_embed
        self.transformer.decoder.class_embed = self.class_embed
        self.transformer.decoder.class_embed[-1] = Still_Classifier(hidden_dim)
        self.transformer.decoder.class_embed[-1].body.bias.data = torch.ones(1) * bias_value
        for box_embed in self.bbox_embed:
            nn.init.constant_(box_embed.layers[-1].bias.data[2:], 0.0)
