"""Multi-GPU plumbing of the hot path; one process per GPU, `torch.distributed` (NCCL on GPUs,
gloo in the CPU tests).

* Volumes are independent: `shard_volumes` deals them round-robin to ranks - no data-path collective (configs 3/4);
  `reduce_validation` turns the per-rank sums of the validation loop into the global averages with one small all-reduce.
* Within ONE volume only the slice encoder shards (config 5): `encode_volume_sharded` lets every rank encode a
  contiguous block of slices and all-gathers the feature pyramid (fpn levels `[32,256,256] + [64,128,128] +
  [256,64,64]` per slice; the position encodings are constant tables and are never sent), filling the predictor's
  per-slice feature cache on every rank.  Propagation stays sequential in t on the rank that tracks the volume.
* Split-KV memory cross-attention: `shard_memory_attention` deals the conditioning memories of
  the bank (and the recent memories) round-robin to the ranks; all ranks run the tracking loop in lockstep, every rank
  attends over its own keys (`ms2_attention_dv_partial`), one all-gather of the 1 MiB per-rank partial (O, m, l) per
  layer is the exchange step, and `ms2_attention_merge` gives every rank the same softmax-normalised result.
"""
import math

import torch
import torch.distributed as dist


def shard_volumes(n_volumes, rank=None, world=None):
    """indices of the volumes rank `rank` tracks (round-robin: equal counts +-1, no exchange needed)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    return list(range(rank, n_volumes, world))


def reduce_validation(loss_sum, metric_sums, n_volumes, group=None):
    """Validation over volumes dealt to ranks by `shard_volumes` (reference: func_3d/function.py:198-314 on one GPU,
    `return tot / n_val, tuple(a / n_val for a in mix_res)`): every rank passes the SUMS over its own volumes of the
    per-volume loss and of the per-volume metric tuple (IoU, Dice, ...) plus its volume count; ONE all-reduce of
    len(metric_sums) + 2 doubles gives every rank the same global averages.  Without an initialised process group
    this is the single-process statement.  Returns (mean loss, tuple of mean metrics)."""
    vec = torch.tensor([float(loss_sum), float(n_volumes), *[float(m) for m in metric_sums]], dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            vec = vec.cuda()
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
        vec = vec.cpu()
    n = float(vec[1])
    if n == 0:
        raise ZeroDivisionError("reduce_validation: no volume was evaluated on any rank")
    return float(vec[0]) / n, tuple(float(v) / n for v in vec[2:])


def slice_block(num_frames, rank, world):
    """contiguous block [lo, hi) of slices encoded by `rank`; blocks have equal length `per` except the tail."""
    per = int(math.ceil(num_frames / world))
    lo = min(rank * per, num_frames)
    return lo, min(lo + per, num_frames), per


@torch.no_grad()
def encode_volume_sharded(predictor, inference_state, group=None):
    """Encode all slices of the volume held by `inference_state`, each rank a contiguous block, then all-gather the
    pyramid so that every rank's `cached_features` holds every slice.  Results equal a local encode with the same slice batches bit for bit
    (a different batching may change the split-KV factor of the global-attention blocks, i.e. the summation order).  Returns the number of slices this
    rank encoded."""
    st = inference_state
    on = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    T = st["num_frames"]
    if getattr(predictor, "feature_cache_size", T) < T:
        raise ValueError(f"feature_cache_size ({predictor.feature_cache_size}) must hold all {T} slices")
    lo, hi, per = slice_block(T, rank, world)
    nb = max(1, int(getattr(predictor, "feature_encode_batch", 1)))
    imgs = st["images"]
    dev = st["device"]
    levels, pos, kept_images = None, None, {}
    for f0 in range(lo, hi, nb):
        f1 = min(f0 + nb, hi)
        if hasattr(imgs, "frames"):                       # host-resident volume: upload only this rank's block
            batch = imgs.frames(f0, f1)
        else:
            batch = torch.stack([imgs[f] for f in range(f0, f1)]).to(dev)
        if batch.dtype != torch.bfloat16:
            batch = batch.float()
        out = predictor.forward_image(batch)
        fpn = [t.permute(0, 2, 3, 1) for t in out["backbone_fpn"]]          # NHWC memory of the channels-last views
        fpn = [t if t.is_contiguous() else t.contiguous() for t in fpn]
        if levels is None:
            levels = [torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for t in fpn]
            pos = [p[:1] for p in out["vision_pos_enc"]]
        for lvl, t in zip(levels, fpn):
            lvl[f0 - lo: f1 - lo] = t
        for i, f in enumerate(range(f0, f1)):
            kept_images[f] = batch[i: i + 1]
    if levels is None:                                    # this rank has no slices: learn the shapes from slice 0
        out = predictor.forward_image(imgs[0].to(dev).unsqueeze(0))
        fpn = [t.permute(0, 2, 3, 1) for t in out["backbone_fpn"]]
        levels = [torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for t in fpn]
        pos = [p[:1] for p in out["vision_pos_enc"]]
    if world > 1:
        gathered = []
        for lvl in levels:
            parts = [torch.empty_like(lvl) for _ in range(world)]
            dist.all_gather(parts, lvl, group=group)                        # the one exchange step of this path
            gathered.append(parts)
    else:
        gathered = [[lvl] for lvl in levels]
    cache = st["cached_features"]
    cache.clear()
    for f in range(T):
        owner, idx = divmod(f, per)
        fpn = [gathered[l][owner][idx: idx + 1].permute(0, 3, 1, 2) for l in range(len(levels))]
        image = kept_images.get(f)
        if image is None and not hasattr(imgs, "frames"):
            image = imgs[f].to(dev).unsqueeze(0)             # device-resident volume: a view; host-resident: not uploaded
        cache[f] = (image, {"vision_features": fpn[-1], "vision_pos_enc": pos, "backbone_fpn": fpn})
    return hi - lo


class KVShard:
    """Placement of the memory bank's keys over the ranks of `group` and the exchange of the attention partials.

    Exchange (`exchange="p2p"`, default when peer memory is available): every rank owns two gather buffers of `world`
    packed partials + flag words in symmetric (peer-mapped) device memory (`torch.distributed._symmetric_memory`, used
    for allocation and address exchange only).  The kernel that folds a rank's local splits stores the partial straight
    into every rank's buffer over NVLink and publishes a step counter (`ms2_attention_dv_partial_push`); the merge kernel
    spins on the flags (`ms2_attention_merge_wait`).  No collective call, no host synchronisation: one exchange per layer
    per frame = (world-1) x 1.03 MiB of peer writes per rank.  `exchange="nccl"`: one all_gather_into_tensor per layer."""

    def __init__(self, rank, world, group=None, exchange=None):
        import os
        self.rank, self.world, self.group = int(rank), int(world), group
        self.exchanges = 0
        self.exchange = exchange or os.environ.get("MS2_KV_EXCHANGE", "p2p")
        self._symm = None
        self.step = 0
        self.nvlink_bytes = 0

    def owns_cond(self, i):
        """conditioning memory number i (arrival order) lives on rank i mod world (rank 0 always owns the first)."""
        return i % self.world == self.rank

    def owns_recent(self, j, n_cond):
        """recent memory j of the current frame; the rotation by the number of conditioning memories evens the load."""
        return (j + n_cond) % self.world == self.rank

    def owns_pointers(self, n_recent, n_cond):
        """the object-pointer tokens (<= 64 rows) travel with the first recent memory, or with the first conditioning
        memory when there is none, so that no rank ends up with a few rows only."""
        return self.owns_recent(0, n_cond) if n_recent else self.rank == 0

    def _symmetric(self, part_numel, device):
        """two gather buffers [2, world, part_numel] fp32 + flags [2, world] int32 in peer-mapped memory; None when the
        platform cannot provide it (falls back to NCCL)."""
        if self._symm is not None and self._symm["numel"] == part_numel:
            return self._symm
        if torch.device(device).type != "cuda":          # gloo / CPU host-logic tests
            self.exchange = "nccl"
            return None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            grp = self.group if self.group is not None else dist.group.WORLD
            buf = symm_mem.empty((2 * self.world * part_numel,), dtype=torch.float32, device=device)
            flg = symm_mem.empty((64,), dtype=torch.int32, device=device)
            flg.zero_()
            hb = symm_mem.rendezvous(buf, grp)
            hf = symm_mem.rendezvous(flg, grp)
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            self._symm = {"numel": part_numel, "buf": buf, "flags": flg, "buf_ptrs": list(hb.buffer_ptrs),
                          "flag_ptrs": list(hf.buffer_ptrs), "counter": torch.zeros(1, dtype=torch.int32, device=device),
                          "handles": (hb, hf)}
        except Exception as e:                          # no peer access / symmetric memory on this platform
            import warnings
            warnings.warn(f"KVShard: peer-memory exchange unavailable ({e}); using NCCL all-gather")
            self.exchange = "nccl"
            self._symm = None
        return self._symm

    def attend(self, q, k, v):
        """q [B,L,256] (identical on all ranks), k/v: this rank's keys / raw values (None or Lk = 0: no share)
        -> softmax(q K_all^T) V_all [B,L,64], identical on all ranks."""
        from . import ops
        B, L, _ = q.shape
        part_numel = B * L * 66
        if self.world > 1 and self.exchange == "p2p":
            sm = self._symmetric(part_numel, q.device)
            if sm is not None:
                self.step += 1
                b = self.step & 1
                dst = [sm["buf_ptrs"][r] + ((b * self.world + self.rank) * part_numel) * 4 for r in range(self.world)]
                flg = [sm["flag_ptrs"][r] + (b * self.world + self.rank) * 4 for r in range(self.world)]
                ops.attention_dv_partial_push(q, k, v, dst, flg, self.step, sm["counter"])
                parts = sm["buf"].view(2, self.world, part_numel)[b]
                flags = sm["flags"][b * self.world: (b + 1) * self.world]
                self.exchanges += 1
                self.nvlink_bytes += (self.world - 1) * part_numel * 4
                return ops.attention_merge_wait(parts, flags, self.step, B, L)
        part = ops.attention_dv_partial(q, k, v)
        if self.world == 1:
            parts = part[None]
        else:
            parts = torch.empty((self.world, part.numel()), dtype=part.dtype, device=part.device)
            dist.all_gather_into_tensor(parts.view(-1), part, group=self.group)      # the exchange step
            self.exchanges += 1
        return ops.attention_merge(parts, B, L)


@torch.no_grad()
def add_prompts_sharded(predictor, inference_state, prompts, group=None):
    """Prompt phase of ONE long volume on several GPUs (config 5).  `prompts`: [(frame_idx, obj_id, bbox), ...] of a single
    object, as func_3d/function.py:236-267 issues them.  The prompted frames are independent of each other (no memory
    is attended to on an initial conditioning frame), so instead of every rank running all of them in lockstep, rank r
    runs the prompt step (prompt encoder + mask decoder) only for the frames whose conditioning memory it will own
    (`KVShard.owns_cond`: the i-th prompted frame, in frame order, lives on rank i mod world); ONE all-gather of the
    low-resolution logits (256 KiB) and object pointers (1 KiB) per frame gives every rank the same per-frame outputs.
    The memory encoder of a prompted frame then also runs on its owner only (`propagate_in_video_preflight`): nobody else
    ever reads that frame's memory features.  Needs `shard_memory_attention(predictor)`; without a process group it
    is the plain loop."""
    st = inference_state
    on = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    shard = predictor.memory_attention.kv_shard
    if world == 1 or shard is None:
        for f, obj_id, bbox in prompts:
            predictor.train_add_new_bbox(inference_state=st, frame_idx=f, obj_id=obj_id, bbox=torch.as_tensor(bbox),
                                         clear_old_points=False)
        return
    obj_ids = {o for _, o, _ in prompts}
    if len(obj_ids) != 1 or st["obj_ids"] and set(st["obj_ids"]) != obj_ids:
        raise ValueError("add_prompts_sharded handles the prompts of ONE object on a fresh state")
    obj_id = next(iter(obj_ids))
    order = sorted(prompts, key=lambda p: p[0])
    if len({p[0] for p in order}) != len(order):
        raise ValueError("add_prompts_sharded: one prompt per frame")
    owner = {p[0]: i % world for i, p in enumerate(order)}          # == KVShard.owns_cond(i) of the consolidated bank
    per = (len(order) + world - 1) // world
    dev = st["device"]
    lo = predictor.image_size // 4
    pm = torch.zeros((per, 1, 1, lo, lo), dtype=torch.float32, device=dev)
    ptr = torch.zeros((per, 1, predictor.hidden_dim), dtype=torch.float32, device=dev)
    slot = {}
    for i, (f, _, bbox) in enumerate(order):
        slot[f] = i // world
        if owner[f] != rank:
            continue
        predictor.train_add_new_bbox(inference_state=st, frame_idx=f, obj_id=obj_id, bbox=torch.as_tensor(bbox),
                                     clear_old_points=False)
        out = st["temp_output_dict_per_obj"][predictor._obj_id_to_idx(st, obj_id)]["cond_frame_outputs"][f]
        pm[slot[f]] = out["pred_masks"]
        ptr[slot[f]] = out["obj_ptr"]
    pm_all = torch.empty((world,) + tuple(pm.shape), dtype=pm.dtype, device=dev)
    ptr_all = torch.empty((world,) + tuple(ptr.shape), dtype=ptr.dtype, device=dev)
    dist.all_gather_into_tensor(pm_all.view(-1), pm.view(-1), group=group)          # the exchange step of the prompt phase
    dist.all_gather_into_tensor(ptr_all.view(-1), ptr.view(-1), group=group)
    obj_idx = predictor._obj_id_to_idx(st, obj_id)
    for f, _, bbox in order:
        if owner[f] == rank:
            continue
        # the bookkeeping `_add_new_points` does, without the compute
        pts, lab = predictor._bbox_points(torch.as_tensor(bbox))
        pts = pts / torch.tensor([st["video_width"], st["video_height"]]).to(pts.device)
        st["point_inputs_per_obj"][obj_idx][f] = {"point_coords": (pts * predictor.image_size).to(dev),
                                                  "point_labels": lab.to(dev)}
        st["mask_inputs_per_obj"][obj_idx].pop(f, None)
        st["temp_output_dict_per_obj"][obj_idx]["cond_frame_outputs"][f] = {
            "maskmem_features": None, "maskmem_pos_enc": None, "pred_masks": pm_all[owner[f], slot[f]],
            "obj_ptr": ptr_all[owner[f], slot[f]]}
    st["_ms2_prompt_owner"] = owner
    st["_ms2_rank"] = rank


def shard_memory_attention(predictor, group=None):
    """Switch `predictor.memory_attention` to split-KV over the ranks of `group` (all ranks must then drive the
    predictor with the same calls).  Returns the KVShard (None when not distributed)."""
    on = dist.is_available() and dist.is_initialized()
    if not on or dist.get_world_size(group) == 1:
        predictor.memory_attention.kv_shard = None
        return None
    shard = KVShard(dist.get_rank(group), dist.get_world_size(group), group)
    predictor.memory_attention.kv_shard = shard
    return shard
