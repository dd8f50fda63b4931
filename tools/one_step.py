"""One warm-up + N profiled hot-path steps at the bench workload (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hipie_b200 import ops
from hipie_b200.modeling import params as P
from hipie_b200.modeling.hipie_img import HIPIE_IMG
prec = 1 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else 3
ops.set_precision(prec)
hp = bench.vit_h_hp(bench.CONFIGS[1])
model = HIPIE_IMG(hp=hp, state_dict=P.random_state_dict(hp, seed=0), device="cuda:0")
model.engine.bf16_value_map = prec == 1
B = 8
dev = torch.device("cuda:0")
imgs = torch.rand(B, 3, 1024, 1024, device=dev) * 255
ids, am, pos_map, is_thing = bench.synth_text(80, 512)
ids_d, am_d = ids.unsqueeze(0).repeat(B, 1).to(dev), am.unsqueeze(0).repeat(B, 1).to(dev)
pad = torch.zeros(B, 1024, 1024, dtype=torch.bool, device=dev)
with torch.no_grad():
    for it in range(2):
        torch.cuda.synchronize()
        if it == 1:
            torch.cuda.cudart().cudaProfilerStart()
        lang = model.engine.forward_text(ids_d, am_d, same_rows=True)
        out = model.coco_inference(imgs, pad, [(1024, 1024)] * B, lang)
        torch.cuda.synchronize()
        if it == 1:
            torch.cuda.cudart().cudaProfilerStop()
print("ok")
