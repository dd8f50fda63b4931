"""CPU: the pybind drop-in `hipie_b200.MultiScaleDeformableAttention` under the UNMODIFIED reference wrapper
(/root/reference/projects/HIPIE/hipie/models/deformable_detr/ops/functions/ms_deform_attn_func.py).  The reference tree only
exists in the build container (not on the GPU box), so here the check is the binding: the module imports with the shim
installed under the pybind name, `MSDeformAttnFunction.apply` reaches `ms_deform_attn_forward` with the reference's positional
argument list, and CPU tensors fail the way the reference op does ("Not implemented on the CPU", ops/src/ms_deform_attn.h:38).
The numerical check of the same call chain on CUDA is tests/test_model_gpu.py::test_pybind_shim_through_reference_style_function."""
import importlib.util
import os
import sys

import pytest
import torch

REF = "/root/reference/projects/HIPIE/hipie/models/deformable_detr/ops/functions/ms_deform_attn_func.py"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present on this box")
def test_reference_wrapper_binds_to_shim(monkeypatch):
    import hipie_b200.MultiScaleDeformableAttention as shim
    monkeypatch.setitem(sys.modules, "MultiScaleDeformableAttention", shim)
    spec = importlib.util.spec_from_file_location("ref_ms_deform_attn_func", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.MSDA is shim
    seen = {}
    real = shim.ms_deform_attn_forward

    def spy(*args):
        seen["n"] = len(args)
        seen["im2col_step"] = args[-1]
        return real(*args)
    monkeypatch.setattr(shim, "ms_deform_attn_forward", spy)
    shapes = torch.tensor([(4, 4), (2, 2)], dtype=torch.long)
    lsi = torch.tensor([0, 16])
    value = torch.rand(1, 20, 2, 4)
    loc = torch.rand(1, 3, 2, 2, 2, 2)
    w = torch.rand(1, 3, 2, 2, 2)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        mod.MSDeformAttnFunction.apply(value, shapes, lsi, loc, w, 64)
    assert seen == {"n": 6, "im2col_step": 64}
    # and the reference's own pure-PyTorch core agrees with the oracle restatement used everywhere else
    from hipie_oracle.msda import ms_deform_attn_core
    assert torch.allclose(mod.ms_deform_attn_core_pytorch(value, shapes, loc, w), ms_deform_attn_core(value, shapes, loc, w))
