"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: volume sharding and the sharded slice encoding with
its all-gather of the feature pyramid.  The encoder itself is a CUDA-only kernel path, so a
deterministic stand-in with the same output structure is used - what is under test is the partition / exchange /
cache-fill plumbing, which must reproduce the single-process cache exactly."""
import os
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from medsam2_b200.parallel import encode_volume_sharded, shard_volumes, slice_block


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
