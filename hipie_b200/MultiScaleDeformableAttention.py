"""Drop-in for the reference's pybind module `MultiScaleDeformableAttention`
(/root/reference/projects/HIPIE/hipie/models/deformable_detr/ops/src/vision.cpp:13-16): same function name and
argument list, so `ops/functions/ms_deform_attn_func.py` works unmodified when this module is importable under
that name (e.g. `sys.modules["MultiScaleDeformableAttention"] = hipie_b200.MultiScaleDeformableAttention`).
Forward only (inference); CPU tensors raise like the reference ("Not implemented on the CPU")."""
import torch

from . import ops


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    for t, n in ((value, "value"), (spatial_shapes, "spatial_shapes"), (level_start_index, "level_start_index"),
                 (sampling_loc, "sampling_loc"), (attn_weight, "attn_weight")):
        if not t.is_contiguous():
            raise RuntimeError(f"{n} tensor has to be contiguous")
        if not t.is_cuda:
            raise RuntimeError(f"{n} must be a CUDA tensor")
    batch = value.shape[0]
    step = min(batch, im2col_step)
    if step > 0 and batch % step != 0:
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")
    return ops.msda_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)


def ms_deform_attn_backward(*args, **kwargs):
    raise RuntimeError("hipie_b200 implements the inference hot path only: ms_deform_attn_backward is out of scope")
