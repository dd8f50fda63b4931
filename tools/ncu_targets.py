"""Launches each hot kernel once at a realistic shape (for ncu captures; not a benchmark)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hipie_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "all"
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
ops.set_precision(prec)
torch.manual_seed(0)
if which in ("all", "attn"):
    B, H, hd, T = 2, 16, 80, 4096
    E = H * hd
    qk = torch.randn(B * T, 2 * E, device=dev)
    v = torch.randn(E, B * T, device=dev)
    S, Vs = ops.split(qk), ops.split(v)
    q, k = ops.BF2(S.hi[:, :E], S.lo[:, :E]), ops.BF2(S.hi[:, E:], S.lo[:, E:])
    rel_h = torch.randn(B, H, T, 64, device=dev)
    rel_w = torch.randn(B, H, T, 64, device=dev)
    for _ in range(2):
        ops.attention_tc(q, k, Vs, B, H, T, hd, T * 2 * E, 2 * E, T * 2 * E, 2 * E, hd ** -0.5, rel_h=rel_h, rel_w=rel_w, kh=64, kw=64)
if which in ("all", "attn16"):      # the single-pass fp16 mode of the precision map (B = 8: one full global block)
    B, H, hd, T = 8, 16, 80, 4096
    E = H * hd
    qk16 = torch.randn(B * T, 2 * E, device=dev).half()
    vt16 = torch.randn(E, B * T, device=dev).half()
    q, k, vt = ops.BF2(qk16[:, :E], None), ops.BF2(qk16[:, E:], None), ops.BF2(vt16, None)
    rel_h = torch.randn(B, H, T, 64, device=dev)
    rel_w = torch.randn(B, H, T, 64, device=dev)
    for _ in range(2):
        ops.attention_tc(q, k, vt, B, H, T, hd, T * 2 * E, 2 * E, T * 2 * E, 2 * E, hd ** -0.5, rel_h=rel_h, rel_w=rel_w, kh=64, kw=64, f16=True)
if which in ("all", "gemm"):
    a = torch.randn(32768, 1280, device=dev)
    w = torch.randn(5120, 1280, device=dev) * 0.02
    A, W = ops.split(a), ops.split_weight(w)
    for _ in range(2):
        ops.gemm(A, W, act=ops.ACT_GELU, want_f32=False, want_split=True)
if which in ("all", "msda"):
    B = 8
    shapes = torch.tensor([(128, 128), (64, 64), (32, 32), (16, 16)], device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    Sx = int(shapes.prod(1).sum())
    value = torch.randn(B, Sx, 256, device=dev)
    packed = torch.cat([torch.randn(B, Sx, 256, device=dev) * 2.0, torch.randn(B, Sx, 128, device=dev)], -1)
    refs = []
    for (h, w_) in shapes.tolist():
        ys, xs = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, device=dev) / h, torch.linspace(0.5, w_ - 0.5, w_, device=dev) / w_, indexing="ij")
        refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    refp = torch.cat(refs, 0)[None, :, None, :].repeat(B, 1, 4, 1).contiguous()
    import hipie_b200.ops as O
    O.MSDA_WINDOWS = True
    for _ in range(2):      # encoder call: shared-memory window kernel (shapes known on the host)
        ops.msda_fused(value, shapes, lsi, packed, refp, shapes_host=[(128, 128), (64, 64), (32, 32), (16, 16)])
if which in ("msda_flat",):      # the same encoder call on the flat L1-gather kernel (A/B of the window kernel)
    import hipie_b200.ops as O
    O.MSDA_WINDOWS = False
    B = 8
    hw = [(128, 128), (64, 64), (32, 32), (16, 16)]
    shapes = torch.tensor(hw, device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    Sx = int(shapes.prod(1).sum())
    value = torch.randn(B, Sx, 256, device=dev)
    packed = torch.cat([torch.randn(B, Sx, 256, device=dev) * 2.0, torch.randn(B, Sx, 128, device=dev)], -1)
    refs = []
    for (h, w_) in hw:
        ys, xs = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, device=dev) / h, torch.linspace(0.5, w_ - 0.5, w_, device=dev) / w_, indexing="ij")
        refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    refp = torch.cat(refs, 0)[None, :, None, :].repeat(B, 1, 4, 1).contiguous()
    for _ in range(2):
        ops.msda_fused(value, shapes, lsi, packed, refp)
if which in ("all", "maskembed"):
    B, HW, Q, C = 8, 65536, 300, 256
    Fm = ops.split(torch.randn(B * HW, C, device=dev))
    Em = ops.split(torch.randn(B * Q, C, device=dev) * 0.1)
    for _ in range(2):
        ops.gemm(Em, Fm, M=Q, N=HW, K=C, batch=B, lda=C, ldw=C, a_bstride=Q * C, w_bstride=HW * C, bits_threshold=0.0)
if which in ("ffn1",):      # deformable-encoder FFN1: K=256, epilogue / store heavy
    a = torch.randn(174080, 256, device=dev)
    w = torch.randn(2048, 256, device=dev) * 0.05
    A, W = ops.split(a), ops.split_weight(w)
    for _ in range(2):
        ops.gemm(A, W, act=ops.ACT_RELU, want_f32=False, want_split=True)
if which in ("proj256",):   # 174080 x 256 x 256 with f32 + split out
    a = torch.randn(174080, 256, device=dev)
    w = torch.randn(256, 256, device=dev) * 0.05
    A, W = ops.split(a), ops.split_weight(w)
    for _ in range(2):
        ops.gemm(A, W, want_f32=True, want_split=True)
if which in ("gemm6",):      # ViT fc1 on the fp16 + e4m3 split (prec 6): fp16 sweep + ONE e4m3 sweep, planes out
    a = torch.randn(32768, 1280, device=dev)
    w = torch.randn(5120, 1280, device=dev) * 0.02
    A, W = ops.split_f16_e4m3(a), ops.split_f16_e4m3(w, weight=True)
    for _ in range(2):
        ops.gemm(A, W, act=ops.ACT_GELU, want_f32=False, out_e4m3=True, prec=6)
if which in ("qk4",):        # ViT qk linear on two fp16 passes (prec 4)
    a = torch.randn(32768, 1280, device=dev)
    w = torch.randn(2560, 1280, device=dev) * 0.02
    A, W = ops.BF2(a.half(), None), ops.split_weight_f16(w)
    for _ in range(2):
        ops.gemm(A, W, want_f32=False, want_split=True, out_fp16=True, prec=4)
if which in ("segpost",):    # semantic + panoptic post-processing of one 1024^2 image: 900 kept + 300 background queries, 80 classes
    Q, C = 1200, 80
    masks = torch.randn(Q, 256, 256, device=dev) * 3.0
    cls = torch.softmax(torch.randn(Q, C, device=dev) * 3.0, -1)
    for _ in range(2):
        ops.seg_postprocess(masks, cls, 0.25, 1024, 1024)
torch.cuda.synchronize()
print("done", which)
