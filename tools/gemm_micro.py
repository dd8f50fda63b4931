"""GEMM micro-benchmark over the hot-path shapes (steering tool): ms and algorithmic TFLOP/s per shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipie_b200 import ops
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ops.set_precision(prec)
dev = torch.device("cuda:0")
torch.manual_seed(0)
SHAPES = [  # name, M, N, K, kwargs
    ("vit fc1 gelu split", 32768, 5120, 1280, dict(act=ops.ACT_GELU, want_f32=False, want_split=True)),
    ("vit fc2 f32", 32768, 1280, 5120, dict(want_f32=True, want_split=False)),
    ("vit qkv split", 32768, 2560, 1280, dict(want_f32=False, want_split=True)),
    ("vit proj f32", 32768, 1280, 1280, dict(want_f32=True, want_split=False)),
    ("vit v^T split", 32768, 1280, 1280, dict(want_f32=False, want_split=True, transposed=True)),
    ("enc ffn1 relu split", 174080, 2048, 256, dict(act=ops.ACT_RELU, want_f32=False, want_split=True)),
    ("enc ffn2 f32", 174080, 256, 2048, dict(want_f32=True, want_split=False)),
    ("enc proj f32+split", 174080, 256, 256, dict(want_f32=True, want_split=True)),
    ("enc offs f32", 174080, 384, 256, dict(want_f32=True, want_split=False)),
    ("dec small", 2400, 256, 256, dict(want_f32=True, want_split=True)),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, M, N, K, kw in SHAPES:
    A = ops.split(torch.randn(M, K, device=dev))
    W = ops.split_weight(torch.randn(N, K, device=dev) * 0.05)
    for _ in range(2):
        ops.gemm(A, W, **kw)
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.gemm(A, W, **kw); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:22s} {M:7d}x{N:5d}x{K:5d}  {ms*1000:8.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TF")
    del A, W
