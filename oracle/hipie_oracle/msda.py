"""Oracle for multi-scale deformable attention (test infrastructure only).

Restates /root/reference/projects/HIPIE/hipie/models/deformable_detr/ops/functions/ms_deform_attn_func.py:43-63
(`ms_deform_attn_core_pytorch`, the grid_sample formulation the reference's own ops/test.py checks its
CUDA kernel against) and, independently, the scalar arithmetic of the reference CUDA kernel
(.../ops/src/cuda/ms_deform_im2col_cuda.cuh:33-84,237-299) as a plain-Python loop for tiny shapes.
"""
import math

import torch
import torch.nn.functional as F


def ms_deform_attn_core(value, value_spatial_shapes, sampling_locations, attention_weights):
    """value (N,S,M,D); spatial shapes [(H,W)...]; loc (N,Lq,M,L,P,2) in [0,1]; weights (N,Lq,M,L,P)
    -> (N, Lq, M*D).   ms_deform_attn_func.py:43-63."""
    N_, S_, M_, D_ = value.shape
    _, Lq_, M_, L_, P_, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in value_spatial_shapes]
    value_list = value.split([h * w for h, w in shapes], dim=1)
    sampling_grids = 2 * sampling_locations - 1
    sampling_value_list = []
    for lid_, (H_, W_) in enumerate(shapes):
        value_l_ = value_list[lid_].flatten(2).transpose(1, 2).reshape(N_ * M_, D_, H_, W_)
        sampling_grid_l_ = sampling_grids[:, :, :, lid_].transpose(1, 2).flatten(0, 1)
        sampling_value_l_ = F.grid_sample(value_l_, sampling_grid_l_, mode="bilinear", padding_mode="zeros",
                                          align_corners=False)
        sampling_value_list.append(sampling_value_l_)
    attention_weights = attention_weights.transpose(1, 2).reshape(N_ * M_, 1, Lq_, L_ * P_)
    output = (torch.stack(sampling_value_list, dim=-2).flatten(-2) * attention_weights).sum(-1).view(N_, M_ * D_, Lq_)
    return output.transpose(1, 2).contiguous()


def ms_deform_attn_scalar(value, value_spatial_shapes, level_start_index, sampling_locations, attention_weights):
    """Pure-Python restatement of the CUDA im2col kernel (ms_deform_im2col_cuda.cuh:237-299 with the
    bilinear helper :33-84).  Tiny shapes only."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    out = torch.zeros(N, Lq, M * D, dtype=value.dtype)
    for b in range(N):
        for q in range(Lq):
            for m in range(M):
                for d in range(D):
                    col = 0.0
                    for l in range(L):
                        H, W = int(value_spatial_shapes[l][0]), int(value_spatial_shapes[l][1])
                        start = int(level_start_index[l])
                        for p in range(P):
                            loc_w = float(sampling_locations[b, q, m, l, p, 0])
                            loc_h = float(sampling_locations[b, q, m, l, p, 1])
                            weight = float(attention_weights[b, q, m, l, p])
                            h_im = loc_h * H - 0.5
                            w_im = loc_w * W - 0.5
                            if h_im > -1 and w_im > -1 and h_im < H and w_im < W:
                                h_low, w_low = math.floor(h_im), math.floor(w_im)
                                h_high, w_high = h_low + 1, w_low + 1
                                lh, lw = h_im - h_low, w_im - w_low
                                hh, hw = 1 - lh, 1 - lw
                                v1 = v2 = v3 = v4 = 0.0
                                if h_low >= 0 and w_low >= 0:
                                    v1 = float(value[b, start + h_low * W + w_low, m, d])
                                if h_low >= 0 and w_high <= W - 1:
                                    v2 = float(value[b, start + h_low * W + w_high, m, d])
                                if h_high <= H - 1 and w_low >= 0:
                                    v3 = float(value[b, start + h_high * W + w_low, m, d])
                                if h_high <= H - 1 and w_high <= W - 1:
                                    v4 = float(value[b, start + h_high * W + w_high, m, d])
                                col += (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4) * weight
                    out[b, q, m * D + d] = col
    return out
