"""CondInst dynamic-mask kernel at the bench shape (8 images x 910 queries, 128^2 coarse -> 256^2 logits): ms and GB/s of output."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipie_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, Q, Hf, Wf = 8, 910, 128, 128
feats = torch.randn(B, Hf, Wf, 8, device=dev)
params = torch.randn(B, Q, 169, device=dev) * 0.3
ref = torch.rand(B, Q, 2, device=dev) * 1000
for _ in range(2):
    out = ops.condinst_masks(feats, params, ref, Hf, Wf)
ts = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); out = ops.condinst_masks(feats, params, ref, Hf, Wf); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ms = sorted(ts)[2]
print(f"condinst {B}x{Q} queries: {ms:.3f} ms, {out.numel() * 4 / ms / 1e6:.0f} GB/s of output, checksum {float(out.double().abs().mean()):.6f}")
