"""LayerNorm at the ViT-H and deformable-encoder shapes: us and GB/s (read + written bytes) per output format."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipie_b200 import ops, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
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


for rows, C in ((32768, 1280), (174080, 256)):
    x = torch.randn(rows, C, device=dev)
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    for name, kw, nbytes in (("bf16 hi/lo", dict(), rows * C * 8), ("fp16", dict(out_fp16=True), rows * C * 6), ("fp16+e4m3", dict(out_e4m3=True), rows * C * 8),
                             ("f32 + bf16 hi/lo", dict(want_f32=True), rows * C * 12)):
        res = []
        for bulk, cap in ((0, 0), (0, 1), (1, 1)):      # round-1 launch shape | grid-stride over resident CTAs | rows staged by bulk copies
            _lib.set_option("ln_bulk", bulk)
            _lib.set_option("ln_grid_cap", cap)
            res.append(timed(lambda: ops.layernorm(x, g, b, 1e-6, **kw)))
        _lib.set_option("ln_bulk", 1)
        _lib.set_option("ln_grid_cap", 1)
        print(f"layernorm {rows}x{C} -> {name:18s} per-8-rows CTAs {res[0]*1000:6.1f} us {nbytes / res[0] / 1e6:5.0f} GB/s | grid-stride {res[1]*1000:6.1f} us "
              f"{nbytes / res[1] / 1e6:5.0f} GB/s | bulk-staged rows {res[2]*1000:6.1f} us {nbytes / res[2] / 1e6:5.0f} GB/s", flush=True)
