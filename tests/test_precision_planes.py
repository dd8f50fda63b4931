"""CPU: the operand-plane formats of the precision map, restated in plain torch.

The device kernels (csrc/common.cuh: split4_f16_e4m3, e4m3_slot0; csrc/gemm_tc.cu prec 6) are checked against these formulas in
tests/test_kernels_gpu.py; here the formulas themselves are pinned: the layout of the e4m3 planes, the reconstruction the host
helper BF2.float() performs, and the accuracy class of the fp16 + e4m3 split product against an fp64 product."""
import math

import pytest
import torch

from hipie_b200.ops import BF2


def f8(t):
    return t.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()


def split_planes(x, weight=False):
    """(h16 (rows, K) fp16, p8 (rows, 2K) uint8): slot 0 / slot 1 interleaved in 32-column groups, byte (k // 32) * 64 + k % 32 (+ 32)."""
    h = x.half()
    lo = x - h.float()
    s0, s1 = (f8(lo * 2.0 ** 14), f8(h.float() * 2.0 ** 4)) if weight else (f8(h.float()), f8(lo * 2.0 ** 10))
    rows, K = x.shape
    p8 = torch.stack([s0.view(rows, K // 32, 32), s1.view(rows, K // 32, 32)], 2).reshape(rows, 2 * K)
    return h, p8.to(torch.float8_e4m3fn).view(torch.uint8)


def test_plane_layout_and_reconstruction():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(7, 96, generator=g) * torch.logspace(-2, 1.5, 96)
    h, p8 = split_planes(x)
    assert p8.shape == (7, 192) and p8.dtype == torch.uint8
    k = 40                                                   # column 40 lives in group 1: bytes 64 + 8 (slot 0) and 64 + 32 + 8 (slot 1)
    v = p8.view(torch.float8_e4m3fn).float()
    assert torch.equal(v[:, 64 + 8], f8(h.float())[:, k]) and torch.equal(v[:, 64 + 32 + 8], f8((x - h.float()) * 1024.0)[:, k])
    rec = BF2(h, p8).float()                                 # hi + slot 1 / 2^10
    assert (rec - x).abs().max() <= (x.abs() * 2.0 ** -15).max()
    assert (h.float() - x).abs().max() > (rec - x).abs().max() * 8      # the second slot carries ~4 more bits


@pytest.mark.parametrize("K", [256, 1280])
def test_split_product_accuracy_class(K):
    """Ah.Wh + 2^-14 (A8 . W8^T) against the fp64 product: the error class of the three-pass bf16 split (2^-16 per product),
    an order of magnitude below one fp16 pass."""
    g = torch.Generator().manual_seed(K)
    a = torch.randn(64, K, generator=g, dtype=torch.float64).float() * 1.5
    w = torch.randn(48, K, generator=g, dtype=torch.float64).float() * 0.05
    ah, a8 = split_planes(a)
    wh, w8 = split_planes(w, weight=True)
    e = lambda p: p.view(torch.float8_e4m3fn).double()
    split = ah.double() @ wh.double().t() + (e(a8) @ e(w8).t()) * 2.0 ** -14
    exact = a.double() @ w.double().t()
    one_pass = ah.double() @ wh.double().t()
    bf = lambda t: (t.bfloat16().double(), (t - t.bfloat16().float()).bfloat16().double())
    (a1, a2), (w1, w2) = bf(a), bf(w)
    bf16x3 = a1 @ w1.t() + a1 @ w2.t() + a2 @ w1.t()
    err_split, err_one, err_x3 = (split - exact).abs().max().item(), (one_pass - exact).abs().max().item(), (bf16x3 - exact).abs().max().item()
    assert err_split < 6e-5 * math.sqrt(K) * 1.5 * 0.05 * 4
    assert err_split < err_one / 8 and err_split < 4 * err_x3 + 1e-7, (err_split, err_one, err_x3)
