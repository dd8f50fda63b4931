"""Per-shape GEMM time of one eager hot step (B=8, 1024^2, ViT-H): which contractions are far from the tensor roofline."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hipie_b200 import ops
from hipie_b200.modeling import params as P
from hipie_b200.modeling.hipie_img import HIPIE_IMG

hp = bench.vit_h_hp(bench.CONFIGS[1])
dev = torch.device("cuda:0")
ops.set_precision(3)
model = HIPIE_IMG(hp=hp, state_dict=P.random_state_dict(hp, seed=0), device="cuda:0")
B = 8
imgs = torch.rand(B, 3, 1024, 1024, device=dev) * 255
ids, am, pos_map, is_thing = bench.synth_text(80, 512)
ids_d, am_d = ids.unsqueeze(0).repeat(B, 1).to(dev), am.unsqueeze(0).repeat(B, 1).to(dev)
pad = torch.zeros(B, 1024, 1024, dtype=torch.bool, device=dev)
sizes = [(1024, 1024)] * B
with torch.no_grad():
    for it in range(3):
        if it == 2:
            ops.profiler.shapes = True
            ops.profiler.start()
        lang = model.engine.forward_text(ids_d, am_d, same_rows=True)
        model.coco_inference(imgs, pad, sizes, lang, task="detection")
    prof = ops.profiler.stop()
rows = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
tot = sum(v["ms"] for v in prof.values())
print(f"total timed {tot:.1f} ms")
for tag, v in rows[:60]:
    per = v["ms"] / v["launches"]
    tf = v["work"] / v["launches"] / per / 1e9 if v["work"] else 0
    print(f"{v['ms']:8.2f} ms {v['ms']/tot*100:5.1f}%  n={v['launches']:4d} avg {per*1000:8.1f} us  {tf:7.0f} (TF alg | GB/s)  {tag}")
