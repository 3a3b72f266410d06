"""GPU idle time inside one bench step: union of all kernel / memcpy intervals (CUPTI timestamps) against the step span,
and the largest gaps with the kernels around them.  usage: python tools/gpu_idle.py [slices]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "medical-sam2_b200"))
import torch
import medsam2_b200
from oracle.config import get_config
from oracle.weights import param_spec
from synth_data import btcv_volume, seeded_weights
from torch.profiler import ProfilerActivity, profile

T = int(sys.argv[1]) if len(sys.argv) > 1 else 96
m = medsam2_b200.build_sam2_video_predictor("sam2_hiera_s", device="cuda", hydra_overrides_extra=[
    "++model.image_size=1024", f"++model.feature_cache_size={T}", "++model.feature_encode_batch=8", "++model.use_cuda_graphs=true",
    "++model.feature_prefetch=true"])
m.load_state_dict(seeded_weights(param_spec(get_config("sam2_hiera_s"))), strict=True)
vol, boxes = btcv_volume(T, 1024, 1234, 1)
vol = vol.cuda()


def run():
    st = m.val_init_state(imgs_tensor=vol, video_height=1024, video_width=1024)
    for f in range(0, T, 2):
        m.train_add_new_bbox(inference_state=st, frame_idx=f, obj_id=1, bbox=torch.tensor(boxes[f][0]), clear_old_points=False)
    for _ in m.propagate_in_video(st, start_frame_idx=0):
        pass


for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run()
    torch.cuda.synchronize()
ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
             if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda x: x[0])
t0, t1 = ev[0][0], max(e[1] for e in ev)
busy, cur_end, gaps, prev = 0.0, t0, [], None
for s, e, n in ev:
    if s > cur_end:
        gaps.append((s - cur_end, prev, n, cur_end - t0))
        busy += 0
        cur_start = s
    busy += max(0.0, e - max(s, cur_end))
    if e > cur_end:
        cur_end, prev = e, n
span = t1 - t0
print(f"# {len(ev)} device activities, span {span / 1e3:.2f} ms, busy (union) {busy / 1e3:.2f} ms, idle {100 * (1 - busy / span):.1f} %")
gaps.sort(reverse=True)
print("# largest gaps: us, at ms, after -> before")
for g, a, b, at in gaps[:25]:
    print(f"{g:8.1f} {at / 1e3:8.2f}  {str(a)[:60]:60s} -> {str(b)[:60]}")
import collections
hist = collections.Counter()
for g, *_ in gaps:
    hist[min(int(g // 5) * 5, 100)] += g
print("# idle time by gap size (us bucket: total ms):", {k: round(v / 1e3, 2) for k, v in sorted(hist.items())})
