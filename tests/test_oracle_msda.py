"""CPU: pin the MSDeformAttn oracle against fixtures generated from the reference's own
ms_deform_attn_core_pytorch (oracle/gen_golden.py; shapes/seed of the reference ops/test.py)."""
import os

import torch

from hipie_oracle.msda import ms_deform_attn_core, ms_deform_attn_scalar


def _cases(golden_dir):
    return torch.load(os.path.join(golden_dir, "msda_core.pt"))


def test_core_matches_reference_fixtures(golden_dir):
    for name, c in _cases(golden_dir).items():
        out = ms_deform_attn_core(c["value"], c["shapes"], c["loc"], c["w"])
        assert out.shape == c["out"].shape
        assert torch.equal(out, c["out"]) or torch.allclose(out, c["out"], rtol=0, atol=1e-7), name


def test_scalar_restatement_of_cuda_kernel_matches_grid_sample(golden_dir):
    # the CUDA kernel arithmetic (im2col + bilinear) == the grid_sample formulation: reference ops/test.py
    # tolerances: fp64 allclose default, fp32 rtol 1e-2 atol 1e-3
    cs = _cases(golden_dir)
    for name in ("test_py_double", "test_py_float"):
        c = cs[name]
        lsi = torch.cat((c["shapes"].new_zeros((1,)), c["shapes"].prod(1).cumsum(0)[:-1]))
        out = ms_deform_attn_scalar(c["value"], c["shapes"], lsi, c["loc"], c["w"])
        if name.endswith("double"):
            assert torch.allclose(out, c["out"])
        else:
            assert torch.allclose(out, c["out"], rtol=1e-2, atol=1e-3)


def test_out_of_range_and_empty():
    shapes = torch.as_tensor([(3, 4)])
    value = torch.ones(1, 12, 1, 2)
    loc = torch.tensor([[-0.5, -0.5], [0.5, 0.5], [1.5, 0.5], [0.0, 0.0]]).view(1, 1, 1, 1, 4, 2)
    w = torch.full((1, 1, 1, 1, 4), 0.25)
    out = ms_deform_attn_core(value, shapes, loc, w)
    # sample 0 and 2 fall outside (zero padding), sample 1 is interior (1.0), sample 3 is the corner (0.25)
    assert torch.allclose(out, torch.full((1, 1, 2), 0.25 * 1.0 + 0.25 * 0.25))
    lsi = torch.zeros(1, dtype=torch.long)
    assert torch.allclose(ms_deform_attn_scalar(value, shapes, lsi, loc, w), out)
    empty = ms_deform_attn_core(value, shapes, loc[:, :0], w[:, :0])
    assert empty.shape == (1, 0, 2)
