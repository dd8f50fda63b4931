"""Times one MaskCLIP call (ViT-L/14-336, random weights) at the two sizes the reference uses per image: 900 foreground mask tokens
(hipie_img.py:599) and ~300 kept + 300 background masks at full resolution (hipie_img.py:737)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from hipie_b200 import ops
from hipie_b200.modeling.maskclip import MaskCLIP
from hipie_oracle import clip as oc          # weights only (random init of the restated architecture); not on the timed path
ops.set_precision(3)
dev = torch.device("cuda:0")
torch.manual_seed(0)
sd = oc.CLIP(oc.VIT_L_14_336).state_dict()
mc = MaskCLIP(sd, device=dev)
img = torch.rand(3, 1024, 1024, device=dev)
for name, Q, up, crop in (("fg 900 masks @1/4 res", 900, 1, None), ("pano 600 masks x4 cropped", 600, 4, (1024, 1024)), ("pano 1200 masks", 1200, 4, (1024, 1024))):
    masks = torch.randn(Q, 256, 256, device=dev) * 3
    for _ in range(2):
        mc.get_mask_embed(img, masks, up=up, crop=crop)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        mc.get_mask_embed(img, masks, up=up, crop=crop)
    e.record(); torch.cuda.synchronize()
    T = Q + 577
    flops = 24 * (2.0 * T * 1024 * (3072 + 1024 + 4096 + 4096) + 4.0 * 16 * T * 577 * 64)
    ms = s.elapsed_time(e) / 3
    print(f"{name}: {ms:.2f} ms per call  ({flops / ms / 1e9:.0f} TF algorithmic, {T} tokens)")
