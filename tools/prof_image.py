import os, sys, time, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/medical-sam2_b200")
import numpy as np, torch
import medsam2_b200
from oracle.config import get_config
from oracle.weights import param_spec
from synth_data import fundus_images, seeded_weights
medsam2_b200.set_compute_dtype(torch.bfloat16)
model = medsam2_b200.build_sam2("sam2_hiera_s", device="cuda", hydra_overrides_extra=["++model.image_size=1024"])
model.use_cuda_graphs = True
model.load_state_dict(seeded_weights(param_spec(get_config("sam2_hiera_s"))), strict=True)
pred = medsam2_b200.SAM2ImagePredictor(model)
imgs, pts = fundus_images(4, 1024, 0)
labels = [np.array([1])] * 4
def step():
    pred.set_image_batch(imgs)
    return pred.predict_batch(pts, labels, multimask_output=True, return_logits=True)
for _ in range(4): step()
torch.cuda.synchronize()
for name, fn in (("set_image_batch", lambda: pred.set_image_batch(imgs)), ("predict_batch", lambda: pred.predict_batch(pts, labels, multimask_output=True, return_logits=True))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); print(name, (time.perf_counter() - t0) / 5 * 1e3, "ms")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]
print("device total ms", sum(e.device_time_total for e in prof.key_averages()) / 1e3)
for e in rows: print(f"{e.device_time_total/1e3:8.2f} ms {e.count:5d} {e.key[:90]}")
rows = sorted(prof.key_averages(), key=lambda e: -e.self_cpu_time_total)[:10]
for e in rows: print(f"cpu {e.self_cpu_time_total/1e3:8.2f} ms {e.count:5d} {e.key[:90]}")
