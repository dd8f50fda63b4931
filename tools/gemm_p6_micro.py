"""fp16 + e4m3 split GEMM (prec 6) on the ViT-H MLP shapes against the 3-pass bf16 mode: ms and algorithmic TFLOP/s."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipie_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
SHAPES = [("vit fc1 gelu", 32768, 5120, 1280, dict(act=ops.ACT_GELU, want_f32=False), True),
          ("vit fc2 +res", 32768, 1280, 5120, dict(), False),
          ("vit proj +res", 32768, 1280, 1280, dict(), False),
          ("enc ffn1 relu", 174080, 2048, 256, dict(act=ops.ACT_RELU, want_f32=False), True),
          ("enc ffn2 +res", 174080, 256, 2048, dict(), False)]
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


for name, M, N, K, kw, planes_out in SHAPES:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.05
    res = None if planes_out else torch.randn(M, N, device=dev)
    A3, W3 = ops.split(a), ops.split_weight(w)
    A6, W6 = ops.split_f16_e4m3(a), ops.split_f16_e4m3(w, weight=True)
    if planes_out:
        t3 = timed(lambda: ops.gemm(A3, W3, want_split=True, prec=3, **kw))
        t6 = timed(lambda: ops.gemm(A6, W6, out_e4m3=True, prec=6, **kw))
    else:
        out = torch.empty(M, N, device=dev)
        t3 = timed(lambda: ops.gemm(A3, W3, residual=res, out_f32=out, prec=3, **kw))
        t6 = timed(lambda: ops.gemm(A6, W6, residual=res, out_f32=out, prec=6, **kw))
    fl = 2.0 * M * N * K
    print(f"{name:14s} {M:6d}x{N:5d}x{K:5d}  bf16x3 {t3*1000:7.1f} us {fl/t3/1e9:6.1f} TF   f16+e4m3 {t6*1000:7.1f} us {fl/t6/1e9:6.1f} TF", flush=True)

# which half costs FFN1 (K = 256, store heavy): mainloop format x output format
M, N, K = 174080, 2048, 256
a = torch.randn(M, K, device=dev)
w = torch.randn(N, K, device=dev) * 0.05
A3, W3 = ops.split(a), ops.split_weight(w)
A6, W6 = ops.split_f16_e4m3(a), ops.split_f16_e4m3(w, weight=True)
for name, fn in (("p3 -> bf16 hi/lo", lambda: ops.gemm(A3, W3, act=ops.ACT_RELU, want_f32=False, want_split=True, prec=3)),
                 ("p3 -> fp16+e4m3", lambda: ops.gemm(A3, W3, act=ops.ACT_RELU, want_f32=False, out_e4m3=True, prec=3)),
                 ("p3 -> fp16 only", lambda: ops.gemm(A3, W3, act=ops.ACT_RELU, want_f32=False, want_split=True, out_fp16=True, prec=3)),
                 ("p6 -> bf16 hi/lo", lambda: ops.gemm(A6, W6, act=ops.ACT_RELU, want_f32=False, want_split=True, prec=6)),
                 ("p6 -> fp16+e4m3", lambda: ops.gemm(A6, W6, act=ops.ACT_RELU, want_f32=False, out_e4m3=True, prec=6)),
                 ("p6 -> fp16 only", lambda: ops.gemm(A6, W6, act=ops.ACT_RELU, want_f32=False, want_split=True, out_fp16=True, prec=6))):
    t = timed(fn)
    print(f"ffn1 {name:18s} {t*1000:7.1f} us", flush=True)
