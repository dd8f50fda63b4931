"""Per-shape GEMM / kernel time breakdown of one full-size hot-path step (steering tool, writes gpurun_out/breakdown.json)."""
import json, os, sys, time
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
def step():
    lang = model.forward_text(ids_d, am_d)
    return model.coco_inference(imgs, pad, [(1024, 1024)] * B, lang)
with torch.no_grad():
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
    ops.profiler.shapes = True
    ops.profiler.start()
    step()
    prof = ops.profiler.stop()
tot = sum(v["ms"] for v in prof.values())
print(f"step wall {wall*1000:.1f} ms; timed kernels {tot:.1f} ms")
rows = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
for k, v in rows[:45]:
    rate = v["work"] / (v["ms"] / 1000) / 1e12 if v["work"] else 0
    print(f"{k:46s} n={v['launches']:4d} total {v['ms']:8.2f} ms  avg {v['ms']/v['launches']:7.3f}  {rate:8.1f} T(FLOP|B)/s")
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"wall_ms": wall * 1000, "timed_ms": tot, "kernels": prof}, open(f"gpurun_out/breakdown_p{prec}.json", "w"), indent=1)
