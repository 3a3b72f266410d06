"""Sanity run of the config-3 secondary variant: 13 objects (BTCV organ count), 24 slices, bbox every 2 slices."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "medical-sam2_b200"))
import torch
import medsam2_b200
from oracle.config import get_config
from oracle.weights import param_spec
from synth_data import btcv_volume, seeded_weights
T, S, NOBJ = int(os.environ.get("T", 24)), 1024, int(os.environ.get("NOBJ", 13))
m = medsam2_b200.build_sam2_video_predictor("sam2_hiera_s", device="cuda", hydra_overrides_extra=[
    f"++model.image_size={S}", f"++model.feature_cache_size={T}", "++model.feature_encode_batch=8", "++model.use_cuda_graphs=true"])
m.load_state_dict(seeded_weights(param_spec(get_config("sam2_hiera_s"))), strict=True)
vol, boxes = btcv_volume(T, S, 1234, NOBJ)
vol = vol.cuda()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = m.val_init_state(imgs_tensor=vol, video_height=S, video_width=S)
    for f in range(0, T, 2):
        for o in range(NOBJ):
            m.train_add_new_bbox(inference_state=st, frame_idx=f, obj_id=o + 1, bbox=torch.tensor(boxes[f][o]), clear_old_points=False)
    n = 0
    for f, ids, masks in m.propagate_in_video(st, start_frame_idx=0):
        assert masks.shape == (NOBJ, 1, S, S) and torch.isfinite(masks).all()
        n += 1
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"rep {rep}: {NOBJ} objects, {n} slices in {dt * 1e3:.1f} ms -> {n / dt:.1f} slices/s ({n * NOBJ / dt:.0f} object-slices/s), "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
