"""Oracle: HIPIE_IMG eval forward on CPU in fp32 (test infrastructure only — never imported by product code).

Restates (H = /root/reference/projects/HIPIE/hipie):
  H/hipie_img.py: forward (eval) :314-420, preprocess_image :880-898, forward_text :900-922, inference :537-766,
                  semantic_inference :870-878, panoptic_inference :473-535, convert_grounding_to_od_logits :1025-1052
  H/models/ddetrs_dn.py: DDETRSegmUniDN.__init__ :90-215, coco_inference :801-978, forward_mask_head_train :1006-1069,
                  MaskHeadSmallConv :1581-1689, post_process_maskdino :244-262
  H/models/ddetrs.py: segmentation_postprocess :1029-1076
  H/models/deformable_detr/bert_model.py:10-154 (HF BertModel, >512-token chunking)
  H/util/box_ops.py:17-31, detectron2 ImageList.from_tensors (structures/image_list.py:59-110),
  H/util/misc.py:288-316 nested_tensor_from_tensor_list.
`hp` is a plain dict of hyper-parameters (see hparams.py) so tiny configurations can be built for tests.
"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision.ops as tvops

from . import condinst
from .detr import (MLP, DeformableDETRDINO, DeformableTransformerVLDINO, FeatureResizer, Joiner, PositionEmbeddingSine,
                   agg_lang_feat, inverse_sigmoid)
from .maskdino import MaskDINOHead
from .resnet import ResNet50
from .vit import ViT


def box_cxcywh_to_xyxy(x):
    x_c, y_c, w, h = x.unbind(-1)
    return torch.stack([(x_c - 0.5 * w), (y_c - 0.5 * h), (x_c + 0.5 * w), (y_c + 0.5 * h)], dim=-1)


class MaskHeadSmallConv(nn.Module):
    """ddetrs_dn.py:1581-1689 (fpns=None, use_raft=False)."""

    def __init__(self, dim):
        super().__init__()
        self.lay1 = nn.Conv2d(dim, dim // 4, 3, padding=1)
        self.lay2 = nn.Conv2d(dim // 4, dim // 32, 3, padding=1)
        self.lay3 = nn.Conv2d(dim, dim, 3, padding=1)
        self.lay4 = nn.Conv2d(dim, dim, 3, padding=1)
        self.jia_dcn = nn.Conv2d(dim, dim, 3, padding=1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, a=1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        fused = F.relu(self.lay3(x[-1]))
        fused = x[-2] + F.interpolate(fused, size=x[-2].shape[-2:], mode="nearest")
        fused = F.relu(self.lay4(fused))
        fused = x[-3] + F.interpolate(fused, size=x[-3].shape[-2:], mode="nearest")
        fused = F.relu(self.jia_dcn(fused))
        fused = F.relu(self.lay1(fused))
        return F.relu(self.lay2(fused))


class BertEncoder(nn.Module):
    """bert_model.py:10-154 with a randomly initialised HF BertModel (no checkpoint files are available)."""

    def __init__(self, hp):
        super().__init__()
        from transformers import BertConfig, BertModel
        b = hp["bert"]
        cfg = BertConfig(vocab_size=b["vocab"], hidden_size=b["hidden"], num_hidden_layers=b["layers"],
                         num_attention_heads=b["heads"], intermediate_size=b["inter"], max_position_embeddings=b["max_pos"],
                         hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        self.model = BertModel(cfg, add_pooling_layer=False)

    def forward(self, x, sep=1012):
        input = x["input_ids"]
        mask = x["attention_mask"]
        if input.shape[1] <= 512:
            outputs = self.model(input_ids=input, attention_mask=mask, output_hidden_states=True)
            return {"masks": mask, "hidden": outputs.hidden_states[1:][-1]}
        # ---- >512 tokens: chunk at '.' / EOS boundaries (:68-135) ----
        PAD_VAL, CLS, SEP, EOS = 0, 101, sep, 102
        bs, seq_len = mask.shape
        all_inputs = []
        for bs_i in range(bs):
            input_bs = input[bs_i].clone()
            mask_bs = mask[bs_i]
            begin = 0
            start_src = 0
            while True:
                seps = torch.where((input_bs == sep) | (input_bs == EOS))[0]
                seps = seps[seps < 510]
                if len(seps) == 0:
                    break
                last_sep = seps[-1]
                first_input = input_bs[:last_sep + 1].clone()
                first_input[-1] = EOS
                first_mask = mask_bs[begin:begin + last_sep + 1] if False else mask_bs[:last_sep + 1]
                first_mask_on = torch.where(first_mask == 1)[0]
                l_valid = len(first_input)
                indices = (start_src, start_src + l_valid, begin, begin + l_valid)
                first_mask_out = torch.zeros(512).to(first_input)
                if start_src == 0:
                    pad = torch.zeros((512 - len(first_input))).to(first_input) + PAD_VAL
                    first_input = torch.cat([first_input, pad], dim=0)
                    first_mask_out[first_mask_on] = 1
                else:
                    pad = torch.zeros((512 - len(first_input) - 1)).to(first_input) + PAD_VAL
                    pad[0] = SEP
                    first_input = torch.cat([torch.tensor([CLS]).to(first_input), first_input, pad], dim=0)
                    first_mask_out[first_mask_on + 1] = 1
                    first_mask_out[0] = 1
                all_inputs.append((bs_i, first_input, first_mask_out, indices))
                start_src = 1
                input_bs = input_bs[l_valid:]
                begin += l_valid
        inputs_actual = torch.stack([x[1] for x in all_inputs])
        masks_actual = torch.stack([x[2] for x in all_inputs])
        outputs = self.model(input_ids=inputs_actual, attention_mask=masks_actual, output_hidden_states=True)
        last_hidden_state = outputs.hidden_states[1:][-1]
        final_hidden_state = torch.zeros(bs, seq_len, last_hidden_state.shape[-1])
        for idx, (bs_i, _, _, (s0, s1, t0, t1)) in enumerate(all_inputs):
            final_hidden_state[bs_i][t0:t1] = last_hidden_state[idx][s0:s1]
        return {"masks": mask, "hidden": final_hidden_state}


class DDETRSegmUniDN(nn.Module):
    def __init__(self, detr, hp):
        super().__init__()
        self.detr = detr
        hidden_dim = detr.transformer.d_model
        self.controller = MLP(hidden_dim, hidden_dim, condinst.NUM_GEN_PARAMS, 3)
        for contr in self.controller.layers:
            nn.init.xavier_uniform_(contr.weight)
            nn.init.zeros_(contr.bias)
        self.mask_head = MaskHeadSmallConv(hidden_dim)
        self.resizer = FeatureResizer(hp.get("lang_dim", 768), hidden_dim)   # DYNAMIC_LABEL_ENC (training only)
        self.mask_dino = MaskDINOHead(detr.backbone.num_channels, hp)
        n_mlp = hp.get("md_dec_layers", 9) + 2
        import copy
        self.mask_dino_cls_embed = nn.ModuleList([copy.deepcopy(self.detr.class_embed[0]) for _ in range(n_mlp)])


class HipieOracle(nn.Module):
    def __init__(self, hp):
        super().__init__()
        self.hp = hp
        if hp["backbone"] == "vit":
            bb = ViT(**hp["vit"])
        else:
            bb = ResNet50()
        hidden = hp.get("hidden_dim", 256)
        joiner = Joiner(bb, PositionEmbeddingSine(hidden // 2, offset=-0.5))
        transformer = DeformableTransformerVLDINO(
            d_model=hidden, nhead=8, num_encoder_layers=hp.get("enc_layers", 6), num_decoder_layers=hp.get("dec_layers", 6),
            dim_feedforward=hp.get("dim_ff", 2048), num_feature_levels=4, two_stage_num_proposals=hp.get("num_queries", 900),
            num_bg=hp.get("num_bg", 10), vl_hidden=hp.get("vl_hidden", 2048), lang_dim=hp.get("lang_dim", 768))
        detr = DeformableDETRDINO(joiner, transformer, 4, lang_dim=hp.get("lang_dim", 768))
        self.detr = DDETRSegmUniDN(detr, hp)
        self.text_encoder = nn.Sequential(OrderedDict([("body", BertEncoder(hp))]))
        self.register_buffer("pixel_mean", torch.tensor([123.675, 116.280, 103.530]).view(3, 1, 1), persistent=False)
        self.register_buffer("pixel_std", torch.tensor([58.395, 57.120, 57.375]).view(3, 1, 1), persistent=False)
        self.num_bg, self.num_fg = hp.get("num_bg", 10), hp.get("num_queries", 900)
        self.mask_stride, self.mask_thres = 4, 0.5
        self.pano_temp, self.object_mask_threshold, self.overlap_threshold = 0.06, 0.25, 0.8
        self.max_pool, self.use_bg_for_pano, self.bg_cls_agnostic = hp.get("max_pool", False), False, hp.get("bg_cls_agnostic", False)
        self.clip = None                 # MaskCLIP re-scoring (MODEL.CLIP.ENABLED, hipie_img.py:249-262): attach_clip()

    def attach_clip(self, maskclip, train_labels, alpha=0.35, beta=0.7, fg_a=0.3, fg_b=1.7, agg_mode="MUL", pano_temp_fg=0.06):
        """hipie_img.py:249-262: `maskclip` = hipie_oracle.clip.MaskCLIPOracle; `train_labels` = the COCO-panoptic prompt-engineered
        label list the reference reads at :72 (list of {id, name})."""
        object.__setattr__(self, "clip", maskclip)          # not a submodule: CLIP is never part of the state_dict (clip.py:125)
        self.train_labels, self.clip_alpha, self.clip_beta = train_labels, alpha, beta
        self.clip_fg_a, self.clip_fg_b, self.clip_agg_mode, self.pano_temp_fg = fg_a, fg_b, agg_mode, pano_temp_fg

    # ---------------------------------------------------------------- stages
    def preprocess(self, images):
        """hipie_img.py:880-898 + misc.py:288-316: normalise, zero-pad to the batch max rounded to the backbone's
        size_divisibility, padding mask True on padded pixels."""
        imgs = [(x - self.pixel_mean) / self.pixel_std for x in images]
        sizes = [tuple(x.shape[-2:]) for x in imgs]
        div = getattr(self.detr.detr.backbone[0].backbone, "size_divisibility", 32)
        H, W = max(s[0] for s in sizes), max(s[1] for s in sizes)
        if div > 1:
            H, W = (H + div - 1) // div * div, (W + div - 1) // div * div
        tensor = torch.zeros(len(imgs), 3, H, W)
        mask = torch.ones(len(imgs), H, W, dtype=torch.bool)
        for i, im in enumerate(imgs):
            tensor[i, :, :im.shape[1], :im.shape[2]] = im
            mask[i, :im.shape[1], :im.shape[2]] = False
        return tensor, mask, sizes

    def forward_text(self, input_ids, attention_mask):
        return self.text_encoder[0]({"input_ids": input_ids, "attention_mask": attention_mask}, sep=1012)

    @torch.no_grad()
    def coco_inference(self, tensor, mask, image_sizes, lang, task="detection", forced=None):
        """ddetrs_dn.py:801-978.  `forced` (dict of index tensors) pins the discontinuous top-k selections."""
        forced = forced or {}
        d = self.detr.detr
        features, pos = d.backbone(tensor, mask)
        if task == "grounding":
            lang_feat_pool = agg_lang_feat(lang["hidden"], lang["masks"]).unsqueeze(1)     # pre-fusion pooling (:809-811)
        srcs, masks, poses = [], [], []
        for l, (src, m) in enumerate(features):
            srcs.append(d.input_proj[l](src))
            masks.append(m)
            poses.append(pos[l])
        src = d.input_proj[3](features[-1][0])
        m = F.interpolate(masks[0][None].float(), size=src.shape[-2:]).to(torch.bool)[0]
        srcs.append(src)
        masks.append(m)
        poses.append(d.backbone[1](src, m).to(src.dtype))
        spatial_shapes = [tuple(s.shape[-2:]) for s in srcs]
        lang = {"hidden": lang["hidden"], "masks": lang["masks"]}
        hs, memory, init_reference, inter_references, lang, aux = d.transformer(srcs, masks, poses, lang,
                                                                                forced_topk=forced.get("topk_fg"))
        out = {"aux": aux, "memory": memory, "srcs": srcs, "features": {k: v[0] for k, v in zip(("res3", "res4", "res5"), features)}}
        # decoupled MaskDINO branch (:863-888)
        feats_md = {k: v[0] for k, v in zip(("res3", "res4", "res5"), features)}
        md = self.detr.mask_dino(feats_md, forced_topk=forced.get("topk_md"))
        lang_for_md = lang_feat_pool if task == "grounding" else lang["hidden"]
        md_logits = self.detr.mask_dino_cls_embed[-1](md["pred_logits"], lang_for_md)       # post_process_maskdino idx=-1
        lvl = hs.shape[0] - 1
        reference = inverse_sigmoid(inter_references[lvl - 1])
        if task == "grounding":
            outputs_class = d.class_embed[lvl](hs[lvl], lang_feat_pool)
        else:
            outputs_class = d.class_embed[lvl](hs[lvl], lang["hidden"])
        tmp = d.bbox_embed[lvl](hs[lvl]) + reference
        out["pred_logits"] = outputs_class
        out["pred_boxes"] = tmp.sigmoid()
        out["pred_boxious"] = d.iou_head[lvl](hs[lvl])
        out["hs"] = hs
        out["inter_references"] = inter_references
        # CondInst masks (:952-973)
        ref_points = inter_references[-2, :, :, :2]
        params = self.detr.controller(hs[lvl])
        ref_px = torch.stack([ref_points[i] * torch.tensor([float(image_sizes[i][1]), float(image_sizes[i][0])])
                              for i in range(len(image_sizes))])
        bs, _, c = memory.shape
        encod, idx = [], 0
        for (h, w) in spatial_shapes[:3]:
            encod.append(memory[:, idx:idx + h * w, :].reshape(bs, h, w, c).permute(0, 3, 1, 2))
            idx += h * w
        decod = self.detr.mask_head(encod)                                               # (B, 8, H/8, W/8)
        out["mask_head_feats"] = decod
        out["pred_masks"] = condinst.dynamic_mask_with_coords(decod, ref_px, params, stride=8).unsqueeze(2)
        out["pred_masks_maskdino"] = md["pred_masks"]
        out["pred_logits_maskdino"] = md_logits
        out["pred_boxes_maskdino"] = md["pred_boxes"]
        out["md"] = md
        out["lang_hidden_fused"] = lang["hidden"]
        return out

    # ---------------------------------------------------------------- post-processing
    @staticmethod
    def convert_grounding_to_od_logits(logits, num_classes, positive_map, is_thing, mode=None, max_pool=False):
        """hipie_img.py:1025-1052"""
        scores = torch.zeros(logits.shape[0], logits.shape[1], num_classes)
        for label_j in positive_map:
            idx = torch.LongTensor(positive_map[label_j])
            if max_pool:
                scores[:, :, label_j - 1] = logits[:, :, idx].max(-1)[0]
            else:
                scores[:, :, label_j - 1] = logits[:, :, idx].mean(-1)
            if mode == "FG" and (not is_thing.get(label_j, True)):
                scores[:, :, label_j - 1] = -9999.0
            elif mode == "BG" and is_thing.get(label_j, True):
                scores[:, :, label_j - 1] = -9999.0
        return scores

    def semantic_inference(self, mask_cls, mask_pred):
        return torch.einsum("qc,qhw->chw", mask_cls, mask_pred.sigmoid())

    def panoptic_inference(self, mask_cls, mask_pred, is_thing):
        """hipie_img.py:473-535"""
        scores, labels = mask_cls.max(-1)
        mask_pred = mask_pred.sigmoid()
        keep = scores > self.object_mask_threshold
        cur_scores, cur_classes, cur_masks = scores[keep], labels[keep], mask_pred[keep]
        cur_prob_masks = cur_scores.view(-1, 1, 1) * cur_masks
        h, w = cur_masks.shape[-2:]
        panoptic_seg = torch.zeros((h, w), dtype=torch.int32)
        segments_info = []
        current_segment_id = 0
        if cur_masks.shape[0] == 0:
            return panoptic_seg, segments_info
        cur_mask_ids = cur_prob_masks.argmax(0)
        stuff_memory_list = {}
        for k in range(cur_classes.shape[0]):
            pred_class = cur_classes[k].item()
            isthing = is_thing.get(int(pred_class + 1), True)
            mask_area = (cur_mask_ids == k).sum().item()
            original_area = (cur_masks[k] >= 0.5).sum().item()
            mask = (cur_mask_ids == k) & (cur_masks[k] >= 0.5)
            if mask_area > 0 and original_area > 0 and mask.sum().item() > 0:
                if mask_area / original_area < self.overlap_threshold:
                    continue
                if not isthing:
                    if int(pred_class) in stuff_memory_list.keys():
                        panoptic_seg[mask] = stuff_memory_list[int(pred_class)]
                        continue
                    stuff_memory_list[int(pred_class)] = current_segment_id + 1
                current_segment_id += 1
                panoptic_seg[mask] = current_segment_id
                segments_info.append({"id": current_segment_id, "isthing": bool(isthing), "category_id": int(pred_class)})
        return panoptic_seg, segments_info

    @torch.no_grad()
    def inference(self, out, image_sizes, positive_map, num_classes, task, is_thing, sizes, clip_inputs=None):
        """hipie_img.py:537-766 (OTA path, demo_only False, decoupled MaskDINO decoder).  With MaskCLIP attached, `clip_inputs[i]` =
        dict(image=(3,H,W) in 0..1, test_labels=[{id,name}], text_embed=(n_prompts, D)) drives the two re-scoring sites :592-609
        and :735-747."""
        from .clip import get_clip_logits
        max_num_inst = 100 if task == "detection" else 1
        fg = self.num_bg
        box_cls, box_pred = out["pred_logits"][:, fg:], out["pred_boxes"][:, fg:]
        mask_pred, iou_pred = out["pred_masks"][:, fg:], out["pred_boxious"][:, fg:]
        box_cls_bg = out["pred_logits_maskdino"]
        mask_pred_bg = out["pred_masks_maskdino"].unsqueeze(2)
        results = []
        for i in range(len(image_sizes)):
            image_size = image_sizes[i]
            has_thing = any(is_thing[i].values())
            logits_per_image = self.convert_grounding_to_od_logits(box_cls[i].unsqueeze(0), num_classes, positive_map, is_thing[i],
                                                                   mode="FG" if has_thing else None, max_pool=self.max_pool)[0]
            if self.clip is not None:
                ci = clip_inputs[i]
                is_thing_mask = ~(logits_per_image[:1] == -9999.0)
                p_model = F.softmax(logits_per_image.sigmoid() / self.pano_temp_fg, dim=-1) if logits_per_image.shape[-1] > 1 else logits_per_image.sigmoid()
                prob = get_clip_logits(self.clip, ci["test_labels"], self.train_labels, mask_pred[i][None, :, 0], ci["image"][None], p_model,
                                       ci["text_embed"], self.clip_alpha, self.clip_beta, self.clip_agg_mode).sigmoid() * is_thing_mask.float()
                prob = torch.sqrt((prob ** self.clip_fg_a) * (iou_pred[i].sigmoid() ** self.clip_fg_b))
            else:
                prob = torch.sqrt(logits_per_image.sigmoid() * iou_pred[i].sigmoid())
            nms_scores, idxs = torch.max(prob, 1)
            boxes_before_nms = box_cxcywh_to_xyxy(box_pred[i])
            keep_indices = tvops.batched_nms(boxes_before_nms, nms_scores, idxs, 0.7)
            prob = prob[keep_indices]
            num_inst = min(max_num_inst, prob.numel())
            box_k = box_pred[i][keep_indices]
            mask_k = mask_pred[i][keep_indices]
            topk_values, topk_indexes = torch.topk(prob.view(-1), num_inst, dim=0)
            topk_boxes = torch.div(topk_indexes, logits_per_image.shape[1], rounding_mode="floor")
            labels = topk_indexes % logits_per_image.shape[1]
            box_k, mask_i = box_k[topk_boxes], mask_k[topk_boxes]
            boxes = box_cxcywh_to_xyxy(box_k) * torch.tensor([image_size[1], image_size[0], image_size[1], image_size[0]], dtype=torch.float32)
            N, C, H, W = mask_i.shape
            m = F.interpolate(mask_i, size=(H * self.mask_stride, W * self.mask_stride), mode="bilinear", align_corners=False)
            m = (m.sigmoid() > self.mask_thres)[:, :, :image_size[0], :image_size[1]]
            res = {"pred_boxes": boxes, "pred_masks": m, "scores": topk_values, "pred_classes": labels,
                   "keep_indices": keep_indices, "topk_boxes": topk_boxes}
            sem = pano = None
            if task == "detection":
                mode = None if (self.use_bg_for_pano or self.bg_cls_agnostic) else "BG"
                logits_bg = self.convert_grounding_to_od_logits(box_cls_bg[i].unsqueeze(0), num_classes, positive_map, is_thing[i],
                                                                mode=mode, max_pool=self.max_pool)[0]
                logits_all = torch.cat([logits_per_image[keep_indices], logits_bg], dim=0)
                mask_all = torch.cat([mask_pred[i][keep_indices], mask_pred_bg[i]], dim=0)
                N, C, H, W = mask_all.shape
                logits_all = F.softmax(logits_all.sigmoid() / self.pano_temp, dim=-1)
                mask_all = F.interpolate(mask_all, size=(H * self.mask_stride, W * self.mask_stride), mode="bilinear", align_corners=False)
                mask_all = mask_all[:, :, :image_size[0], :image_size[1]]
                if self.clip is not None:
                    ci = clip_inputs[i]
                    clip_logits = get_clip_logits(self.clip, ci["test_labels"], self.train_labels, mask_all[None, :, 0], ci["image"][None],
                                                  logits_all, ci["text_embed"], self.clip_alpha, self.clip_beta, self.clip_agg_mode)
                    logits_all = clip_logits.softmax(-1)
                mask_up = F.interpolate(mask_all, size=sizes[i], mode="bilinear", align_corners=False)[:, 0]
                sem = self.semantic_inference(logits_all, mask_up)
                pano = self.panoptic_inference(logits_all, mask_up, is_thing[i])
                res["cls_prob_all"] = logits_all
            results.append({"instances": res, "panoptic_seg": pano, "sem_seg": sem})
        return results

    @staticmethod
    def segmentation_postprocess(res, img_size, output_height, output_width):
        """models/ddetrs.py:1029-1076: rescale+clip boxes, drop empty, nearest-resize masks."""
        sx, sy = output_width / img_size[1], output_height / img_size[0]
        boxes = res["pred_boxes"].clone()
        boxes[:, 0::2] *= sx
        boxes[:, 1::2] *= sy
        boxes[:, 0::2] = boxes[:, 0::2].clamp(0, output_width)
        boxes[:, 1::2] = boxes[:, 1::2].clamp(0, output_height)
        keep = ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)
        masks = F.interpolate(res["pred_masks"].float(), size=(output_height, output_width), mode="nearest")[:, 0].to(torch.uint8)
        return {"pred_boxes": boxes[keep], "pred_masks": masks[keep], "scores": res["scores"][keep],
                "pred_classes": res["pred_classes"][keep]}

    @torch.no_grad()
    def forward(self, batched_inputs, input_ids, attention_mask, forced=None):
        """batched_inputs: list of dicts {image (3,H,W) float 0..255, height, width, task, is_thing,
        positive_map_label_to_token}; token ids are synthesised (no tokenizer vocab in this environment)."""
        task = batched_inputs[0]["task"]
        tensor, mask, image_sizes = self.preprocess([x["image"] for x in batched_inputs])
        positive_map = {1: [0]} if task == "grounding" else batched_inputs[0]["positive_map_label_to_token"]
        num_classes = len(positive_map)
        lang = self.forward_text(input_ids, attention_mask)
        out = self.coco_inference(tensor, mask, image_sizes, lang, task=task, forced=forced)
        is_thing = [x["is_thing"] for x in batched_inputs]
        sizes = [(x.get("height", s[0]), x.get("width", s[1])) for x, s in zip(batched_inputs, image_sizes)]
        clip_inputs = None
        if self.clip is not None:          # hipie_img.py:348-352: the un-normalised image / 255, open_seg_labels of every input
            clip_inputs = [dict(image=x["image"] / 255.0, test_labels=x["open_seg_labels"],
                                text_embed=self.clip.build_text_embed(x["clip_prompt_ids"])) for x in batched_inputs]
        results = self.inference(out, image_sizes, positive_map, num_classes, task, is_thing, sizes, clip_inputs=clip_inputs)
        for r, x, s in zip(results, batched_inputs, image_sizes):
            r["instances_post"] = self.segmentation_postprocess(r["instances"], s, x.get("height", s[0]), x.get("width", s[1]))
        return results, out
