"""MSDeformAttn encoder call (B = 8, 1024^2 levels, Lq = S = 21 760): flat L1-gather kernel vs the shared-memory window kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hipie_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
for (img, B) in ((1024, 8), (1280, 4)):
    hw = [(img // 8, img // 8), (img // 16, img // 16), (img // 32, img // 32), (img // 64, img // 64)]
    shapes = torch.tensor(hw, device=dev)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    value = torch.randn(B, S, 256, device=dev)
    refs = []
    for (h, w_) in hw:
        ys, xs = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, device=dev) / h, torch.linspace(0.5, w_ - 0.5, w_, device=dev) / w_, indexing="ij")
        refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    refp = torch.cat(refs, 0)[None, :, None, :].repeat(B, 1, 4, 1).contiguous()
    alg = B * (S * 256 * 4 + S * 128 * 3 * 4 + S * 256 * 4)
    for sigma in (1.0, 2.0, 4.0):
        packed = torch.cat([torch.randn(B, S, 256, device=dev) * sigma, torch.randn(B, S, 128, device=dev)], -1)
        res = {}
        for name, win in (("flat", False), ("window", True)):
            ops.MSDA_WINDOWS = win
            fn = lambda: ops.msda_fused(value, shapes, lsi, packed, refp, shapes_host=hw)
            for _ in range(3):
                out = fn()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                fn()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 10
            res[name] = (ms, out)
            print(f"{img}^2 B={B} offsets sigma={sigma} px  {name:6s}: {ms:.3f} ms  {alg/ms/1e6:.0f} GB/s algorithmic", flush=True)
        assert torch.equal(res["flat"][1].hi, res["window"][1].hi)
ops.MSDA_WINDOWS = False
