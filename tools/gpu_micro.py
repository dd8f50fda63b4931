"""Micro-benchmarks of the individual kernels at the real ViT-H / DETR shapes (CUDA-event timed).
Writes gpurun_out/micro.json.  Not a bench.py replacement: used to steer optimisation."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipie_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = torch.device("cuda:0")
    res = {}
    B = 8
    T = 4096
    # --- GEMMs of a ViT-H block at batch 8
    for name, (M, N, K) in {"qkv": (B * T, 3840, 1280), "proj": (B * T, 1280, 1280), "fc1": (B * T, 5120, 1280),
                            "fc2": (B * T, 1280, 5120), "ffn_detr": (8 * 21760, 2048, 256)}.items():
        a = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) * 0.02
        A, W = ops.split(a), ops.split_weight(w)
        del a
        out = torch.empty(M, N, device=dev)
        for prec in (1, 3):
            ms = timeit(lambda: ops.gemm(A, W, out_f32=out, prec=prec))
            res[f"gemm_{name}_p{prec}"] = {"ms": ms, "tflops_alg": 2 * M * N * K / ms / 1e9, "tflops_mma": prec * 2 * M * N * K / ms / 1e9}
            print(name, prec, res[f"gemm_{name}_p{prec}"], flush=True)
        del A, W, out
    # cuBLAS reference points
    a = torch.randn(B * T, 1280, device=dev, dtype=torch.bfloat16)
    w = torch.randn(3840, 1280, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: a @ w.t())
    res["cublas_bf16_qkv"] = {"ms": ms, "tflops": 2 * B * T * 1280 * 3840 / ms / 1e9}
    print("cublas", res["cublas_bf16_qkv"], flush=True)
    del a, w
    # --- attention (global, 64x64 grid, 16 heads, hd 80), batch 2 images
    Ba, H, hd = 2, 16, 80
    qkv = torch.randn(Ba, T, 3, H, hd, device=dev)
    S = ops.split(qkv)
    del qkv
    ts, bs = 3 * H * hd, T * 3 * H * hd
    q, k, v = ops.BF2(S.hi[:, :, 0], S.lo[:, :, 0]), ops.BF2(S.hi[:, :, 1], S.lo[:, :, 1]), ops.BF2(S.hi[:, :, 2], S.lo[:, :, 2])
    relh = torch.randn(Ba, H, T, 64, device=dev)
    relw = torch.randn(Ba, H, T, 64, device=dev)
    Rt = torch.randn(64, hd, 64, device=dev)
    flops = 4 * T * T * hd * H * Ba
    for prec in (1, 3):
        ms = timeit(lambda: ops.attention(q, k, v, Ba, H, T, T, hd, (bs, ts, hd), (bs, ts, hd), (bs, ts, hd), hd ** -0.5,
                                          rel_h=relh, rel_w=relw, kh=64, kw=64, prec=prec), iters=5)
        res[f"attn_global_p{prec}"] = {"ms": ms, "tflops_alg": flops / ms / 1e9}
        ms2 = timeit(lambda: ops.attention(q, k, v, Ba, H, T, T, hd, (bs, ts, hd), (bs, ts, hd), (bs, ts, hd), hd ** -0.5,
                                           prec=prec), iters=5)
        res[f"attn_global_norel_p{prec}"] = {"ms": ms2, "tflops_alg": flops / ms2 / 1e9}
        print("attn", prec, ms, ms2, flush=True)
    ms = timeit(lambda: ops.relpos_bias(q, (bs, ts, hd), Rt, 0, 64, 64, Ba, H, hd))
    res["relpos_bias"] = {"ms": ms}
    del S, q, k, v, relh, relw
    # --- MSDA at the encoder shape, batch 8
    shapes = torch.tensor([(128, 128), (64, 64), (32, 32), (16, 16)], device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    Sx = int(shapes.prod(1).sum())
    value = torch.randn(B, Sx, 256, device=dev)
    # reference points = pixel centres, offsets ~ N(0, 2 px) like a trained model
    packed = torch.cat([torch.randn(B, Sx, 256, device=dev) * 2.0, torch.randn(B, Sx, 128, device=dev)], -1)
    refp = torch.rand(B, Sx, 1, 2, device=dev).repeat(1, 1, 4, 1).contiguous()
    # raster-ordered reference points per level (what the encoder uses)
    refs = []
    for (h, w) in shapes.tolist():
        ys, xs = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, device=dev) / h, torch.linspace(0.5, w - 0.5, w, device=dev) / w, indexing="ij")
        refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    refp = torch.cat(refs, 0)[None, :, None, :].repeat(B, 1, 4, 1).contiguous()
    alg_bytes = 4 * (256 * Sx + 640 * Sx) * B
    for vdt in ("f32", "bf16"):
        val = value if vdt == "f32" else value.bfloat16()
        ms = timeit(lambda: ops.msda_fused(val, shapes, lsi, packed, refp, want_split=True))
        res[f"msda_enc_{vdt}"] = {"ms": ms, "alg_GBs": alg_bytes / ms / 1e6}
        print("msda", vdt, res[f"msda_enc_{vdt}"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/micro.json", "w"), indent=1)


if __name__ == "__main__":
    main()
