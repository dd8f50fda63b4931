"""Wall-clock breakdown of the public-API call (HIPIE_IMG.forward) at the bench workload: which stage the e2e time goes to."""
import os, sys, time
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
host_imgs = [(torch.rand(3, 1024, 1024) * 255).pin_memory() for _ in range(B)]
ids, am, pos_map, is_thing = bench.synth_text(80, 512)
T = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t
        return r
    return w
model.enable_cuda_graphs(True)          # the serving mode bench.py measures
for n in ("preprocess_image", "_graphed_hot_path", "fused_sem_pano_launch", "fused_sem_pano_finish", "inference", "segmentation_postprocess", "semantic_inference",
          "panoptic_inference", "convert_grounding_to_od_logits"):
    setattr(model, n, timed(n, getattr(model, n)))
def step():
    batched = [dict(image=im, height=1024, width=1024, task="detection", is_thing=is_thing, positive_map_label_to_token=pos_map,
                    input_ids=ids, attention_mask=am) for im in host_imgs]
    res = model(batched)
    torch.cuda.synchronize(); t = time.perf_counter()
    host = []
    for r in res:
        inst = r["instances"]
        host.append((inst.pred_boxes.tensor.cpu(), inst.scores.cpu(), inst.pred_classes.cpu(), r["panoptic_seg"][0].cpu(),
                     r["sem_seg"].argmax(0).to(torch.uint8).cpu()))
    torch.cuda.synchronize(); T["d2h+argmax"] = T.get("d2h+argmax", 0.0) + time.perf_counter() - t
    return host
with torch.no_grad():
    step(); step(); T.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    N = 3
    for _ in range(N):
        step()
    torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / N
    print(f"total {tot*1e3:.1f} ms / batch of {B}")
    for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
        print(f"  {k:34s} {v / N * 1e3:8.2f} ms")
    if len(sys.argv) > 2:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as pr:
            step()
        print(pr.key_averages().table(sort_by="cuda_time_total", row_limit=25))
