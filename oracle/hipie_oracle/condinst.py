"""Oracle: CondInst dynamic mask head (test infrastructure only).

Restates /root/reference/projects/HIPIE/hipie/models/ddetrs_dn.py: mask_heads_forward :1390-1408,
dynamic_mask_with_coords :1411-1502, parse_dynamic_params :1806-1829, aligned_bilinear :1832-1854,
compute_locations :1857-1870 (rel_coord=True, use_raft=False, mask_out_stride=4, 3 controller layers,
8 dynamic channels, in_channels = 256 // 32 = 8).
"""
import torch
import torch.nn.functional as F

IN_CHANNELS = 8
DYN_CHANNELS = 8
WEIGHT_NUMS = [(IN_CHANNELS + 2) * DYN_CHANNELS, DYN_CHANNELS * DYN_CHANNELS, DYN_CHANNELS * 1]   # ddetrs_dn.py:113-128
BIAS_NUMS = [DYN_CHANNELS, DYN_CHANNELS, 1]
NUM_GEN_PARAMS = sum(WEIGHT_NUMS) + sum(BIAS_NUMS)   # 169


def compute_locations(h, w, stride=1):
    """ddetrs_dn.py:1857-1870"""
    shifts_x = torch.arange(0, w * stride, step=stride, dtype=torch.float32)
    shifts_y = torch.arange(0, h * stride, step=stride, dtype=torch.float32)
    shift_y, shift_x = torch.meshgrid(shifts_y, shifts_x, indexing="ij")
    return torch.stack((shift_x.reshape(-1), shift_y.reshape(-1)), dim=1) + stride // 2


def aligned_bilinear(tensor, factor):
    """ddetrs_dn.py:1832-1854"""
    assert tensor.dim() == 4 and factor >= 1 and int(factor) == factor
    if factor == 1:
        return tensor
    h, w = tensor.size()[2:]
    tensor = F.pad(tensor, pad=(0, 1, 0, 1), mode="replicate")
    oh, ow = factor * h + 1, factor * w + 1
    tensor = F.interpolate(tensor, size=(oh, ow), mode="bilinear", align_corners=True)
    tensor = F.pad(tensor, pad=(factor // 2, 0, factor // 2, 0), mode="replicate")
    return tensor[:, :, :oh - 1, :ow - 1]


def parse_dynamic_params(params, channels=DYN_CHANNELS, weight_nums=WEIGHT_NUMS, bias_nums=BIAS_NUMS):
    """ddetrs_dn.py:1806-1829"""
    assert params.dim() == 2 and params.size(1) == sum(weight_nums) + sum(bias_nums)
    num_insts = params.size(0)
    num_layers = len(weight_nums)
    splits = list(torch.split_with_sizes(params, weight_nums + bias_nums, dim=1))
    ws, bs = splits[:num_layers], splits[num_layers:]
    for l in range(num_layers):
        if l < num_layers - 1:
            ws[l] = ws[l].reshape(num_insts * channels, -1, 1, 1)
            bs[l] = bs[l].reshape(num_insts * channels)
        else:
            ws[l] = ws[l].reshape(num_insts * 1, -1, 1, 1)
            bs[l] = bs[l].reshape(num_insts)
    return ws, bs


def mask_heads_forward(features, weights, biases, num_insts):
    """ddetrs_dn.py:1390-1408"""
    x = features
    n_layers = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = F.conv2d(x, w, bias=b, stride=1, padding=0, groups=num_insts)
        if i < n_layers - 1:
            x = F.relu(x)
    return x


def dynamic_mask_with_coords(mask_feats, ref_px, mask_head_params, stride=8, mask_out_stride=4):
    """mask_feats (B, 8, H, W); ref_px (B, Q, 2) pixels; mask_head_params (B, Q, 169)
    -> (B, Q, 2H, 2W).   ddetrs_dn.py:1411-1502 (rel_coord path), batched over images exactly like the
    reference's per-image loop :1441-1455."""
    B, C, H, W = mask_feats.shape
    Q = ref_px.shape[1]
    locations = compute_locations(H, W, stride=stride)                      # (H*W, 2)
    outs = []
    for b in range(B):
        rel = ref_px[b].reshape(1, Q, 1, 1, 2) - locations.reshape(1, 1, H, W, 2)
        rel = rel.float().permute(0, 1, 4, 2, 3).flatten(-2, -1)           # (1, Q, 2, HW)
        feats_b = mask_feats[b].reshape(1, C, H * W).unsqueeze(1).repeat(1, Q, 1, 1)
        head_in = torch.cat([rel, feats_b], dim=2).reshape(1, -1, H, W)     # (1, Q*(C+2), H, W)
        weights, biases = parse_dynamic_params(mask_head_params[b])
        logits = mask_heads_forward(head_in, weights, biases, Q).reshape(-1, 1, H, W)
        logits = aligned_bilinear(logits, int(stride / mask_out_stride))
        outs.append(logits.reshape(Q, logits.shape[-2], logits.shape[-1]))
    return torch.stack(outs, 0)
