import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipie_b200 import ops, _lib
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ops.set_precision(3 if prec == 2 else prec)
dev = torch.device("cuda:0")
B, H, hd, T = 2, 16, 80, 4096
E = H * hd
qk = torch.randn(B * T, 2 * E, device=dev); v = torch.randn(E, B * T, device=dev)
S, Vs = ops.split(qk), ops.split(v)
if prec == 2:      # single fp16 planes (the default ViT attention mode)
    S, Vs = ops.BF2(qk.half(), None), ops.BF2(v.half(), None)
rel_h = torch.randn(B, H, T, 64, device=dev); rel_w = torch.randn(B, H, T, 64, device=dev)
out = ops._empty_bf2((B, T, E), dev)
trace = torch.zeros(32 * 2 * 8, dtype=torch.int64, device=dev)
lib = _lib.load()
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
lo = prec == 3
for it in range(3):
    rc = lib.hipie_attention_tc_traced(P(S.hi), P(S.lo) if lo else None, T * 2 * E, 2 * E, 0, E, P(S.hi[:, E:]), P(S.lo[:, E:]) if lo else None, T * 2 * E, 2 * E, 0, E,
                                       P(Vs.hi), P(Vs.lo) if lo else None, B * T, P(rel_h), P(rel_w), 64, 64, None, P(out.hi), P(out.lo) if lo else None,
                                       T * E, E, B, H, T, hd, hd ** -0.5, prec, P(trace), None)
    assert rc == 0, lib.hipie_last_error()
torch.cuda.synchronize()
t = trace.cpu().view(32, 2, 8)
t0 = int(t[8, 0, 0])
names = ["sm:start", "sm:S ready", "sm:max done", "sm:P stored", "mm:k ok", "mm:QKi/v ok", "mm:p_full ok", "mm:PV issued"]
print("tile qt " + " ".join(f"{n[:12]:>12s}" for n in names))
for j in range(8, 18):
    for q in range(2):
        print(f"{j:4d} {q:2d} " + " ".join(f"{int(t[j, q, k]) - t0:12d}" for k in range(8)))
