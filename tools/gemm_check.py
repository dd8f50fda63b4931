"""GPU smoke + A/B of the CTA-pair (cta_group::2) GEMM tiles against the single-CTA tiles (run under `timeout`).
Exit code 0 = pair tiles produce the same results as single-CTA tiles on every case and both agree with an fp64 reference."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hipie_b200 import _lib, ops

torch.manual_seed(0)
dev = torch.device("cuda:0")
ok = True


def run(M, N, K, prec, pairs, **kw):
    _lib.set_option("gemm_cta_pairs", pairs)
    return ops.gemm(A, W, prec=prec, **kw)


cases = [(1024, 256, 256, {}), (4096, 1280, 1280, dict(act=ops.ACT_GELU, want_split=True)), (2048 + 64, 384, 512, {}),
         (1500, 130, 264, {}), (8192, 128, 256, dict(want_split=True, want_f32=False)), (4096, 640, 1280, dict(transposed=True, want_split=True, want_f32=False)),
         (32768, 1280, 1280, dict(res=True))]
for prec in (3, 1):
    for M, N, K, kw in cases:
        a = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) * 0.05
        bias = torch.randn(N, device=dev)
        ops.set_precision(prec)
        A, W = ops.split(a), ops.split_weight(w)
        kw = dict(kw)
        res = torch.randn(M, N, device=dev) if kw.pop("res", False) else None
        outs = []
        for pairs in (0, 1):
            f, s, _ = run(M, N, K, prec, pairs, bias=bias, residual=res, **kw)
            torch.cuda.synchronize()
            outs.append(f if f is not None else s.float())
        same = torch.equal(outs[0], outs[1])
        ref = a.double() @ w.double().t() + bias.double()
        if kw.get("act") == ops.ACT_GELU:
            ref = torch.nn.functional.gelu(ref)
        if res is not None:
            ref = ref + res.double()
        if kw.get("transposed"):
            ref = ref.t()
        err = (outs[1].double() - ref).abs().max().item()
        tol = 2e-3 if prec == 3 else 0.5
        good = same and err < tol
        ok &= good
        print(f"prec{prec} {M}x{N}x{K} {kw}: pairs==single {same}, max err vs fp64 {err:.2e} {'ok' if good else 'FAIL'}", flush=True)
ops.set_precision(3)
# timing: ViT-H fc1 (32768 x 5120 x 1280, GELU, planes out) and qk (32768 x 2560 x 1280)
for (M, N, K, kw) in [(32768, 5120, 1280, dict(act=ops.ACT_GELU, want_split=True, want_f32=False)), (32768, 2560, 1280, dict(want_split=True, want_f32=False)),
                      (32768, 1280, 5120, dict()), (174080, 2048, 256, dict(act=ops.ACT_RELU, want_split=True, want_f32=False))]:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.05
    A, W = ops.split(a), ops.split_weight(w)
    for prec in (3, 1):
        for pairs in (0, 1):
            _lib.set_option("gemm_cta_pairs", pairs)
            for _ in range(3):
                ops.gemm(A, W, prec=prec, **kw)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                ops.gemm(A, W, prec=prec, **kw)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 10
            print(f"time prec{prec} pairs={pairs} {M}x{N}x{K}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.0f} TF alg ({2.0*M*N*K*prec/ms/1e9:.0f} TF executed)", flush=True)
_lib.set_option("gemm_cta_pairs", 1)
print("GEMM_CHECK_OK" if ok else "GEMM_CHECK_FAIL")
sys.exit(0 if ok else 1)
