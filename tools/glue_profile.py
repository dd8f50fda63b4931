"""Which python lines launch the torch (non-hipie) kernels of one hot-path step (steering tool)."""
import os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hipie_b200 import ops
from hipie_b200.modeling import params as P
from hipie_b200.modeling.hipie_img import HIPIE_IMG
from torch.profiler import profile, ProfilerActivity
ops.set_precision(3)
hp = bench.vit_h_hp(bench.CONFIGS[1])
model = HIPIE_IMG(hp=hp, state_dict=P.random_state_dict(hp, seed=0), device="cuda:0")
B = 8
dev = torch.device("cuda:0")
imgs = torch.rand(B, 3, 1024, 1024, device=dev) * 255
ids, am, pos_map, is_thing = bench.synth_text(80, 512)
ids_d, am_d = ids.unsqueeze(0).repeat(B, 1).to(dev), am.unsqueeze(0).repeat(B, 1).to(dev)
pad = torch.zeros(B, 1024, 1024, dtype=torch.bool, device=dev)
def step():
    lang = model.forward_text(ids_d, am_d)
    return model.coco_inference(imgs, pad, [(1024, 1024)] * B, lang)
with torch.no_grad():
    step(); step(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as pr:
        step()
        torch.cuda.synchronize()
agg = collections.Counter(); cnt = collections.Counter()
for e in pr.events():
    if e.device_type.name != "CPU" or not e.name.startswith("aten::") or e.self_device_time_total <= 0:
        continue
    key = (e.name, str(e.input_shapes)[:110])
    agg[key] += e.self_device_time_total; cnt[key] += 1
tot = sum(agg.values())
print(f"torch-op device time {tot/1e3:.2f} ms")
for k, v in agg.most_common(40):
    print(f"{v/1e3:7.3f} ms n={cnt[k]:4d} {k[0]:22s} {k[1]}")
