"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: volume sharding and the sharded slice encoding with
its all-gather of the feature pyramid.  The encoder itself is a CUDA-only kernel path, so a
deterministic stand-in with the same output structure is used - what is under test is the partition / exchange /
cache-fill plumbing, which must reproduce the single-process cache exactly."""
import os
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from medsam2_b200.parallel import encode_volume_sharded, reduce_validation, shard_volumes, slice_block


class FakePredictor:
    feature_cache_size = 64
    feature_encode_batch = 3

    def forward_image(self, img):                       # [B,3,H,W] -> pyramid with the product's structure
        B, _, H, W = img.shape
        f0 = torch.cat([img, img.flip(1)], 1)[:, :4].permute(0, 2, 3, 1).contiguous()                 # NHWC memory
        f1 = torch.nn.functional.avg_pool2d(img, 2).repeat(1, 2, 1, 1).permute(0, 2, 3, 1).contiguous()
        f2 = torch.nn.functional.avg_pool2d(img, 4).repeat(1, 3, 1, 1).permute(0, 2, 3, 1).contiguous() * 2.0
        fpn = [t.permute(0, 3, 1, 2) for t in (f0, f1, f2)]                                            # channels-last views
        pos = [torch.ones(1, t.shape[1], t.shape[2], t.shape[3]).expand(B, -1, -1, -1) for t in fpn]
        return {"vision_features": fpn[-1], "vision_pos_enc": pos, "backbone_fpn": fpn}


def _state(T):
    g = torch.Generator().manual_seed(3)
    return {"num_frames": T, "images": torch.randn(T, 3, 16, 16, generator=g), "device": torch.device("cpu"),
            "cached_features": OrderedDict()}


def _worker(rank, world, port, T, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = _state(T)
    n = encode_volume_sharded(FakePredictor(), st)
    ref = _state(T)
    encode_volume_sharded_single(ref)
    ok = len(st["cached_features"]) == T
    for f in range(T):
        a, b = st["cached_features"][f], ref["cached_features"][f]
        ok &= torch.equal(a[0], b[0])
        ok &= all(torch.equal(x, y) and x.stride() == y.stride() for x, y in zip(a[1]["backbone_fpn"], b[1]["backbone_fpn"]))
        ok &= all(torch.equal(x, y) for x, y in zip(a[1]["vision_pos_enc"], b[1]["vision_pos_enc"]))
        ok &= torch.equal(a[1]["vision_features"], b[1]["vision_features"])
    counts = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n]))
    ok &= sum(int(c) for c in counts) == T
    out[rank] = bool(ok)
    dist.destroy_process_group()


def encode_volume_sharded_single(st):
    """single-process statement: every slice encoded locally, one at a time."""
    p = FakePredictor()
    for f in range(st["num_frames"]):
        img = st["images"][f].float().unsqueeze(0)
        st["cached_features"][f] = (img, p.forward_image(img))


def _run(T, world=2, port=29631):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, T, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


def test_sharded_slice_encoding_matches_local_even():
    _run(8, port=29631)


def test_sharded_slice_encoding_matches_local_ragged():
    _run(7, port=29632)          # blocks of 4 + 3: the padded tail of the last block must not leak into the cache


def test_sharded_slice_encoding_fewer_slices_than_ranks():
    _run(1, port=29633)          # rank 1 encodes nothing and still receives slice 0


def test_volume_sharding_partitions():
    for n, w in ((64, 8), (5, 2), (3, 4), (0, 2)):
        parts = [shard_volumes(n, r, w) for r in range(w)]
        assert sorted(v for p in parts for v in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert slice_block(10, 0, 4) == (0, 3, 3) and slice_block(10, 3, 4) == (9, 10, 3)
    assert slice_block(2, 3, 4) == (2, 2, 1)


# ---------------------------------------------------------------- split-KV memory attention over ranks (§8(f) rank 1)
def _mem_attn_inputs(B=1, hw=64, n_cond=5, n_recent=3, n_ptr=4):
    g = torch.Generator().manual_seed(11)
    r = lambda *s: torch.randn(*s, generator=g)
    pos = r(hw, 64) * 0.5
    cond = [(("c", i), None, r(B, hw, 64), pos) for i in range(n_cond)]
    recent = [(("r", j), None, r(B, hw, 64), pos + 0.1 * (j + 1)) for j in range(n_recent)]
    return r(B, hw, 256), r(B, hw, 256) * 0.3, cond, recent, r(B, 4 * n_ptr, 64), torch.zeros(B, 4 * n_ptr, 64)


def _kv_worker(rank, world, port, out):
    import pytest
    import medsam2_b200
    import ref_ops
    from medsam2_b200.modeling.memory_attention import MemoryBank
    from medsam2_b200.parallel import KVShard, shard_memory_attention
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mp_ = pytest.MonkeyPatch()
    ref_ops.install(mp_)
    torch.manual_seed(0)
    ok = True
    with medsam2_b200.compute(torch.bfloat16), torch.no_grad():
        m = medsam2_b200.build_sam2_video_predictor("sam2_hiera_t", device="cpu", hydra_overrides_extra=["++model.image_size=128"])
        ma = m.memory_attention
        curr, curr_pos, cond, recent, ptrs, ptr_pos = _mem_attn_inputs()
        ref1 = ma.forward_tokens_banked(curr, curr_pos, MemoryBank(), cond[:2], recent[:1], ptrs, ptr_pos)
        ref2 = ma.forward_tokens_banked(curr, curr_pos, MemoryBank(), cond, recent, ptrs, ptr_pos)
        ref3 = ma.forward_tokens_banked(curr, curr_pos, MemoryBank(), cond[:1], [], ptrs, ptr_pos)
        shard = shard_memory_attention(m)
        assert isinstance(shard, KVShard) and shard.world == world and ma.kv_shard is shard
        bank = MemoryBank()
        got1 = ma.forward_tokens_banked(curr, curr_pos, bank, cond[:2], recent[:1], ptrs, ptr_pos)
        got2 = ma.forward_tokens_banked(curr, curr_pos, bank, cond, recent, ptrs, ptr_pos)     # bank grows: 2 -> 5 cond
        got3 = ma.forward_tokens_banked(curr, curr_pos, MemoryBank(), cond[:1], [], ptrs, ptr_pos)   # rank 1: no keys at all
        ok &= shard.exchanges == 3 * len(ma.layers)
        # this rank holds only its share of the conditioning memories
        ok &= bank.n_static == 64 * len([i for i in range(5) if i % world == rank])
        for a, b in ((got1, ref1), (got2, ref2), (got3, ref3)):
            ok &= bool(torch.isfinite(a).all()) and (a - b).abs().max().item() <= 6e-2 * max(1.0, b.abs().max().item())
        # all ranks hold the same result (the merge consumes the same gathered partials in the same order)
        theirs = [torch.empty_like(got2) for _ in range(world)]
        dist.all_gather(theirs, got2.contiguous())
        ok &= all(torch.equal(t, theirs[0]) for t in theirs)
    mp_.undo()
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_split_kv_memory_attention_two_ranks():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_kv_worker, args=(2, 29733, out), nprocs=2, join=True)
    assert out[0] and out[1], dict(out)


def _prompt_worker(rank, world, port, out):
    """sharded prompt phase + split-KV tracking on 2 gloo ranks == the single-process run (host logic; torch statements)"""
    import pytest
    import medsam2_b200
    import ref_ops
    from medsam2_b200.parallel import add_prompts_sharded, encode_volume_sharded, shard_memory_attention
    from oracle.config import get_config
    from oracle.weights import make_state_dict
    from synth_data import btcv_volume
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mp_ = pytest.MonkeyPatch()
    ref_ops.install(mp_)
    ok = True
    size, T, prompts_at = 128, 7, (0, 2, 4, 5)
    with medsam2_b200.compute(torch.bfloat16), torch.no_grad():
        m = medsam2_b200.build_sam2_video_predictor("sam2_hiera_t", device="cpu", hydra_overrides_extra=[
            f"++model.image_size={size}", "++model.feature_cache_size=16", "++model.feature_encode_batch=2"])
        m.load_state_dict(make_state_dict(get_config("sam2_hiera_t")), strict=True)
        vol, boxes = btcv_volume(T, size, 5, 1)

        def run(sharded):
            st = m.val_init_state(imgs_tensor=vol, video_height=size, video_width=size)
            st["device"] = st["storage_device"] = torch.device("cpu")
            prompts = [(f, 1, boxes[f][0]) for f in prompts_at]
            if sharded:
                encode_volume_sharded(m, st)
                add_prompts_sharded(m, st, prompts)
            else:
                for f, o, b in prompts:
                    m.train_add_new_bbox(inference_state=st, frame_idx=f, obj_id=o, bbox=torch.tensor(b), clear_old_points=False)
            return {f: mk.clone() for f, _, mk in m.propagate_in_video(st, start_frame_idx=0)}, st
        ref, _ = run(False)
        shard_memory_attention(m)
        got, st = run(True)
        # every rank ran the prompt step / memory encoder only for the prompted frames it owns ...
        mine = [f for i, f in enumerate(prompts_at) if i % world == rank]
        cond = st["output_dict"]["cond_frame_outputs"]
        ok &= sorted(f for f in cond if cond[f]["maskmem_features"] is not None) == mine
        # ... and tracks the volume to the single-process masks (bf16 operands; split-KV changes the summation order)
        for f in range(T):
            a, b = got[f].float(), ref[f].float()
            keep = ((a - 0.1).abs() > 1e-6) & ((b - 0.1).abs() > 1e-6)
            ok &= bool((a - b)[keep].abs().max().item() <= 5e-2)
        mine_t = torch.stack([got[f] for f in range(T)]).contiguous()
        theirs = [torch.empty_like(mine_t) for _ in range(world)]
        dist.all_gather(theirs, mine_t)
        ok &= all(torch.equal(t, theirs[0]) for t in theirs)
    mp_.undo()
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_prompt_phase_two_ranks():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_prompt_worker, args=(2, 29737, out), nprocs=2, join=True)
    assert out[0] and out[1], dict(out)


def test_kv_shard_partition_is_exact():
    """every conditioning / recent memory and the pointer tokens belong to exactly one rank, rank 0 owns the first
    conditioning memory, and the pointer tokens never sit alone on a rank."""
    from medsam2_b200.parallel import KVShard
    for world in (1, 2, 3, 8):
        shards = [KVShard(r, world) for r in range(world)]
        for n_cond in range(1, 20):
            for i in range(n_cond):
                assert sum(s.owns_cond(i) for s in shards) == 1
            assert shards[0].owns_cond(0)
            for n_recent in range(0, 7):
                for j in range(n_recent):
                    assert sum(s.owns_recent(j, n_cond) for s in shards) == 1
                owners = [s for s in shards if s.owns_pointers(n_recent, n_cond)]
                assert len(owners) == 1
                o = owners[0]
                assert o.owns_recent(0, n_cond) if n_recent else o.owns_cond(0)


def _val_worker(rank, world, port, n_volumes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    per_volume = torch.rand(n_volumes, 3, generator=g, dtype=torch.float64)        # (loss, iou, dice) of every volume
    mine = shard_volumes(n_volumes)
    loss, metrics = reduce_validation(per_volume[mine, 0].sum(), per_volume[mine, 1:].sum(0).tolist(), len(mine))
    want = per_volume.mean(0)
    out[rank] = bool(abs(loss - want[0]) < 1e-12 and all(abs(a - b) < 1e-12 for a, b in zip(metrics, want[1:].tolist())))
    dist.destroy_process_group()


def test_validation_reduce_two_ranks():
    """volumes dealt round-robin to 2 ranks (5 volumes: 3 + 2), per-rank sums -> identical global averages on both ranks,
    equal to the single-process means (func_3d/function.py:308-314)"""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = mp.Manager().dict()
    mp.spawn(_val_worker, args=(2, port, 5, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_validation_reduce_single_process():
    loss, metrics = reduce_validation(3.0, [1.5, 0.6], 3)
    import pytest
    assert loss == 1.0 and metrics == pytest.approx((0.5, 0.2), abs=1e-15)
    with pytest.raises(ZeroDivisionError):
        reduce_validation(0.0, [0.0], 0)


def _val_driver_worker(rank, world, port, out):
    """validation_sam(shard=True) on 2 ranks == the single-process call (host logic; native ops -> torch statements)"""
    import pytest
    import ref_ops
    from oracle.config import get_config
    from oracle.weights import make_state_dict
    from synth_data import btcv_volume
    import medsam2_b200
    from medsam2_b200.validation import validation_sam
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(4)
    mpatch = pytest.MonkeyPatch()
    ref_ops.install(mpatch)
    size, T = 512, 2
    m = medsam2_b200.build_sam2_video_predictor("sam2_hiera_t", device="cpu", hydra_overrides_extra=[f"++model.image_size={size}"])
    m.load_state_dict(make_state_dict(get_config("sam2_hiera_t")), strict=True)
    packs = []
    for seed in (1234, 99):
        vol, boxes = btcv_volume(T, size, seed, 1)
        g = torch.Generator().manual_seed(seed)
        packs.append({"image": vol, "label": {f: {1: (torch.rand(1, size, size, generator=g) > 0.5).float()} for f in range(T)},
                      "bbox": {f: {1: torch.tensor(boxes[f][0])} for f in range(T)}})
    packs.append({"image": btcv_volume(T, size, 5, 1)[0], "label": {f: {} for f in range(T)}, "bbox": {}})
    sharded = validation_sam(m, packs, prompt="bbox", prompt_freq=1, shard=True, device="cpu")
    ok = True
    if rank == 0:
        whole = validation_sam(m, packs, prompt="bbox", prompt_freq=1, device="cpu")
        ok = abs(sharded[0] - whole[0]) < 1e-9 and all(abs(a - b) < 1e-12 for a, b in zip(sharded[1], whole[1]))
    both = [None, None]
    dist.all_gather_object(both, sharded)
    ok = ok and both[0] == both[1]                                      # every rank holds the same averages
    out[rank] = bool(ok)
    mpatch.undo()
    dist.destroy_process_group()


def test_validation_driver_sharded_two_ranks():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = mp.Manager().dict()
    mp.spawn(_val_driver_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
