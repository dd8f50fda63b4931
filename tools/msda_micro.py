"""Times the fused MSDeformAttn encoder launch (B=8, S=Lq=21760, 8 heads) with grid-like sampling offsets."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipie_b200 import ops
dev = torch.device("cuda:0")
ops.set_precision(3)
torch.manual_seed(0)
B = 8
shapes = torch.tensor([(128, 128), (64, 64), (32, 32), (16, 16)], device=dev)
lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
Sx = int(shapes.prod(1).sum())
value = torch.randn(B, Sx, 256, device=dev)
spread = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0       # offset std in pixels (reference init: up to +-4 px)
packed = torch.cat([torch.randn(B, Sx, 256, device=dev) * spread, torch.randn(B, Sx, 128, device=dev)], -1)
refs = []
for (h, w_) in shapes.tolist():
    ys, xs = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, device=dev) / h, torch.linspace(0.5, w_ - 0.5, w_, device=dev) / w_, indexing="ij")
    refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
refp = torch.cat(refs, 0)[None, :, None, :].repeat(B, 1, 4, 1).contiguous()
for _ in range(3):
    o = ops.msda_fused(value, shapes, lsi, packed, refp)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    o = ops.msda_fused(value, shapes, lsi, packed, refp)
e.record(); torch.cuda.synchronize()
print(f"tile={os.environ.get('HIPIE_MSDA_TILE', 'default')} spread={spread}: {s.elapsed_time(e) / 10 * 1000:.1f} us; checksum {float(o[0].float().abs().sum() if isinstance(o, tuple) else o.hi.float().abs().sum()):.6e}")
