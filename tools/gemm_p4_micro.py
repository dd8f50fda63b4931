"""Two-pass fp16 GEMM (prec 4) on the qkv shapes against the 3-pass bf16 and the one-pass fp16 modes: ms and algorithmic TFLOP/s."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipie_b200 import ops, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
SHAPES = [("vit qk (global)", 32768, 2560, 1280, {}), ("vit v^T (global)", 32768, 1280, 1280, dict(transposed=True)),
          ("vit qk (windows)", 39200, 2560, 1280, {}), ("vit v^T (windows)", 39200, 1280, 1280, dict(transposed=True))]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(7):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


for name, M, N, K, kw in SHAPES:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.05
    A3, W3 = ops.split(a), ops.split_weight(w)
    A4, W4 = ops.BF2(a.half(), None), ops.split_weight_f16(w)
    out = {}
    out["p3"] = timed(lambda: ops.gemm(A3, W3, want_f32=False, want_split=True, out_fp16=True, prec=3, **kw))
    # (the 32-element k-block variant of the CTA-pair kernel measured 4 % slower -- profiles/r02_gemm_p4_micro.txt -- and was removed)
    out["p4"] = timed(lambda: ops.gemm(A4, W4, want_f32=False, want_split=True, out_fp16=True, prec=4, **kw))
    out["f16x1"] = timed(lambda: ops.gemm(A4, W4, want_f32=False, want_split=True, out_fp16=True, prec=2, **kw))
    print(f"{name:20s} {M:6d}x{N:5d}x{K:5d} " + "  ".join(f"{k} {v*1000:7.1f} us {2.0*M*N*K/v/1e9:6.1f} TF" for k, v in out.items()), flush=True)
