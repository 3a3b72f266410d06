"""CUDA-event breakdown of one config-3 volume by phase (encoder / memory attention / SAM heads / memory encoder / rest)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "medical-sam2_b200"))
import torch
import medsam2_b200
from oracle.config import get_config
from oracle.weights import param_spec
from synth_data import btcv_volume, seeded_weights
T, S = 96, 1024
m = medsam2_b200.build_sam2_video_predictor("sam2_hiera_s", device="cuda", hydra_overrides_extra=[
    f"++model.image_size={S}", f"++model.feature_cache_size={T}", "++model.feature_encode_batch=8", "++model.use_cuda_graphs=true",
    "++model.feature_prefetch=" + os.environ.get("PREFETCH", "false")])
m.load_state_dict(seeded_weights(param_spec(get_config("sam2_hiera_s"))), strict=True)
vol, boxes = btcv_volume(T, S, 1234, 1)
vol = vol.cuda()
rec = collections.defaultdict(list)
def wrap(obj, name, label):
    fn = getattr(obj, name)
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(*a, **k); e1.record(); rec[label].append((e0, e1)); return out
    setattr(obj, name, w)
def run():
    st = m.val_init_state(imgs_tensor=vol, video_height=S, video_width=S)
    for f in range(0, T, 2):
        m.train_add_new_bbox(inference_state=st, frame_idx=f, obj_id=1, bbox=torch.tensor(boxes[f][0]), clear_old_points=False)
    for _ in m.propagate_in_video(st, start_frame_idx=0):
        pass
for _ in range(3):
    run()
wrap(m, "forward_image", "image encoder (12 batches of 8)")
wrap(m.memory_attention, "forward_tokens_banked", "memory attention (48 tracked frames)")
wrap(m, "_forward_sam_heads", "prompt encoder + mask decoder heads")
wrap(m, "_encode_new_memory", "memory encoder")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
acc = 0.0
for k, v in rec.items():
    ms = sum(a.elapsed_time(b) for a, b in v); acc += ms
    print(f"{k:45s} {ms:8.2f} ms  {len(v):4d} calls  {1e3 * ms / len(v):8.1f} us/call  {100 * ms / tot:5.1f}%")
print(f"{'everything else (glue, resize, hole filling ...)':45s} {tot - acc:8.2f} ms {'':27s} {100 * (tot - acc) / tot:5.1f}%")
print(f"{'total':45s} {tot:8.2f} ms -> {1e3 * T / tot:.1f} slices/s")
