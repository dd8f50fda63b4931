"""Times hipie_attention_tc at the ViT-H global-block shape (B=8, 16 heads, 4096 tokens, hd 80)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipie_b200 import ops
dev = torch.device("cuda:0")
B, H, hd, T = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 16, 80, 4096
E = H * hd
torch.manual_seed(0)
qk = torch.randn(B * T, 2 * E, device=dev)
v = torch.randn(E, B * T, device=dev)
S, Vs = ops.split(qk), ops.split(v)
q, k = ops.BF2(S.hi[:, :E], S.lo[:, :E]), ops.BF2(S.hi[:, E:], S.lo[:, E:])
rel_h = torch.randn(B, H, T, 64, device=dev)
rel_w = torch.randn(B, H, T, 64, device=dev)
for prec in (3, 1):
    for _ in range(2):
        ops.attention_tc(q, k, Vs, B, H, T, hd, T * 2 * E, 2 * E, T * 2 * E, 2 * E, hd ** -0.5, rel_h=rel_h, rel_w=rel_w, kh=64, kw=64, prec=prec)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        ops.attention_tc(q, k, Vs, B, H, T, hd, T * 2 * E, 2 * E, T * 2 * E, 2 * E, hd ** -0.5, rel_h=rel_h, rel_w=rel_w, kh=64, kw=64, prec=prec)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f"prec {prec}: {ms:.3f} ms  {4.0*B*H*T*T*hd/ms/1e9:.1f} TF algorithmic")
q16, k16, v16 = ops.BF2(qk.half()[:, :E], None), ops.BF2(qk.half()[:, E:], None), ops.BF2(v.half(), None)
for _ in range(2):
    ops.attention_tc(q16, k16, v16, B, H, T, hd, T * 2 * E, 2 * E, T * 2 * E, 2 * E, hd ** -0.5, rel_h=rel_h, rel_w=rel_w, kh=64, kw=64, f16=True)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    ops.attention_tc(q16, k16, v16, B, H, T, hd, T * 2 * E, 2 * E, T * 2 * E, 2 * E, hd ** -0.5, rel_h=rel_h, rel_w=rel_w, kh=64, kw=64, f16=True)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
print(f"fp16 single pass: {ms:.3f} ms  {4.0*B*H*T*T*hd/ms/1e9:.1f} TF algorithmic")
