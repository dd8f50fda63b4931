"""Times hipie_seg_postprocess at the bench shape (Q kept+bg queries, 80 classes, 256^2 -> 1024^2)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipie_b200 import ops
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
C = int(sys.argv[2]) if len(sys.argv) > 2 else 80
dev = torch.device("cuda:0")
masks = torch.randn(Q, 256, 256, device=dev) * 4
cls = torch.softmax(torch.randn(Q, C, device=dev) * 3, -1)
for _ in range(2):
    ops.seg_postprocess(masks, cls, 0.25, 1024, 1024)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ops.profiler.start()
for _ in range(5):
    ops.seg_postprocess(masks, cls, 0.25, 1024, 1024)
prof = ops.profiler.stop()
for k, v in prof.items():
    print(k, v)
