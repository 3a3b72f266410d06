"""Kernel timeline of ONE tracked frame (CUPTI timestamps): every kernel of memory attention, SAM heads, memory encoder
and glue in launch order with its start offset, duration and the gap to the previous kernel.
usage: python tools/frame_timeline.py [slices]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "medical-sam2_b200"))
import torch
import medsam2_b200
from oracle.config import get_config
from oracle.weights import param_spec
from synth_data import btcv_volume, seeded_weights
from torch.profiler import ProfilerActivity, profile

T = int(sys.argv[1]) if len(sys.argv) > 1 else 48
m = medsam2_b200.build_sam2_video_predictor("sam2_hiera_s", device="cuda", hydra_overrides_extra=[
    "++model.image_size=1024", f"++model.feature_cache_size={T}", "++model.feature_encode_batch=8", "++model.use_cuda_graphs=true",
    "++model.feature_prefetch=false"])
m.load_state_dict(seeded_weights(param_spec(get_config("sam2_hiera_s"))), strict=True)
vol, boxes = btcv_volume(T, 1024, 1234, 1)
vol = vol.cuda()


def run(profile_frame=None):
    st = m.val_init_state(imgs_tensor=vol, video_height=1024, video_width=1024)
    for f in range(0, T, 2):
        m.train_add_new_bbox(inference_state=st, frame_idx=f, obj_id=1, bbox=torch.tensor(boxes[f][0]), clear_old_points=False)
    prof = None
    for f, _, mk in m.propagate_in_video(st, start_frame_idx=0):
        if prof is not None:
            torch.cuda.synchronize()
            prof.__exit__(None, None, None)
            return prof
        if profile_frame is not None and f == profile_frame - 1:
            torch.cuda.synchronize()
            prof = profile(activities=[ProfilerActivity.CUDA])
            prof.__enter__()
    return prof


run(); run()
p = run(profile_frame=T - 1)          # the last tracked frame (odd index): longest memory bank
ev = sorted((e for e in p.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
prev_end = t0
tot = 0.0
print(f"# tracked frame {T - 1} of a {T}-slice volume (Lk ~ {(T // 2 + 6) * 4096}); times in us")
print("#   start     dur     gap   kernel")
for e in ev:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    print(f"{s:9.1f} {d:7.1f} {e.time_range.start - prev_end:7.1f}   {e.name[:100]}")
    prev_end = max(prev_end, e.time_range.end)
    tot += d
print(f"# {len(ev)} kernels/copies, busy {tot:.1f} us, span {prev_end - t0:.1f} us")
