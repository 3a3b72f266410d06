import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, R + "/medical-sam2_b200"); sys.path.insert(0, R + "/tests")
import torch
from oracle.config import get_config
from oracle.weights import make_state_dict, param_spec
from synth_data import btcv_volume, seeded_weights
import medsam2_b200
from medsam2_b200 import runtime
spec = param_spec(get_config("sam2_hiera_t"))
def build(graphs):
    m = medsam2_b200.build_sam2_video_predictor("sam2_hiera_t", device="cuda", hydra_overrides_extra=["++model.image_size=512"])
    m.load_state_dict(make_state_dict(get_config("sam2_hiera_t")), strict=True)
    m.use_cuda_graphs = graphs
    m.feature_cache_size, m.feature_encode_batch = 8, 2
    return m
def run(m, prefetch=None):
    m.feature_prefetch = prefetch
    vol, boxes = btcv_volume(8, 512, 99, 1)
    st = m.val_init_state(imgs_tensor=vol.cuda(), video_height=512, video_width=512)
    for f in (0, 4):
        m.train_add_new_bbox(inference_state=st, frame_idx=f, obj_id=1, bbox=torch.tensor(boxes[f][0]), clear_old_points=False)
    return {f: mk.clone() for f, _, mk in m.propagate_in_video(st, start_frame_idx=0)}
r = build(False); r.load_state_dict(seeded_weights(spec, seed=5), strict=True)
c = run(r)
for graphs in (False, True):
    for prefetch in (False, None):
        m = build(graphs)
        a = run(m, prefetch)
        g0 = runtime.generation()
        m.load_state_dict(seeded_weights(spec, seed=5), strict=True)
        b = run(m, prefetch)
        print("graphs", graphs, "prefetch", prefetch, "gen", g0, runtime.generation(), "replays", m._graphs.replays,
              [bool(torch.equal(b[f], c[f])) for f in range(8)], [round((b[f]-c[f]).abs().max().item(), 4) for f in range(8)])
        b2 = run(m, prefetch)
        print("   second run after load:", [bool(torch.equal(b2[f], c[f])) for f in range(8)])
