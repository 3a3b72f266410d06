"""Frame ingest and mask post-processing helpers with the reference's names (utils/misc.py)."""
import os

import numpy as np
import torch

from .. import ops


def get_sdpa_settings():
    """Kept for signature compatibility (utils/misc.py:17-44): there is no SDPA backend choice here."""
    return False, False, True


def get_connected_components(mask):
    """utils/misc.py:47-63: (N,1,H,W) binary mask -> (labels, counts)."""
    from .. import _C
    return _C.get_connected_componnets(mask.to(torch.uint8).contiguous())


def fill_holes_in_mask_scores(mask, max_area):
    """utils/misc.py:247-258 as ONE fused kernel: background components of (mask<=0) with area<=max_area
    are set to +0.1; labels/areas never leave shared memory."""
    assert max_area > 0, "max_area must be positive"
    m = mask.float().contiguous()
    H, W = m.shape[-2:]
    if (H // 2) * (W // 2) <= 28672 and H % 2 == 0 and W % 2 == 0:
        return ops.fill_holes(m, max_area)
    labels, areas = get_connected_components(m <= 0)
    is_hole = (labels > 0) & (areas <= max_area)
    return torch.where(is_hole, 0.1, m)


def mask_to_box(masks):
    """utils/misc.py:66-89 (not on the per-slice path; plain tensor code)."""
    B, _, h, w = masks.shape
    device = masks.device
    xs = torch.arange(w, device=device, dtype=torch.int32)
    ys = torch.arange(h, device=device, dtype=torch.int32)
    gx, gy = torch.meshgrid(xs, ys, indexing="xy")
    gx = gx[None, None].expand(B, 1, h, w)
    gy = gy[None, None].expand(B, 1, h, w)
    min_xs, _ = torch.min(torch.where(masks, gx, w).flatten(-2), dim=-1)
    max_xs, _ = torch.max(torch.where(masks, gx, -1).flatten(-2), dim=-1)
    min_ys, _ = torch.min(torch.where(masks, gy, h).flatten(-2), dim=-1)
    max_ys, _ = torch.max(torch.where(masks, gy, -1).flatten(-2), dim=-1)
    return torch.stack((min_xs, min_ys, max_xs, max_ys), dim=-1)


_UPLOAD_STREAMS = {}


def _upload_stream(device):
    """one long-lived copy stream per device (a fresh stream per volume would give the caching allocator a fresh,
    empty pool every time: cudaMalloc in the timed path)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    s = _UPLOAD_STREAMS.get(key)
    if s is None:
        s = _UPLOAD_STREAMS[key] = torch.cuda.Stream(device=key)
    return s


def _frame_dtype(device_frames=True):
    """dtype of DEVICE-resident normalised frames: bf16 when the GEMM operands are bf16 (the patch-embed contraction
    rounds the frame to bf16 anyway, exactly what autocast does in the reference's train_3d.py:28,57 setting; storing
    it that way halves the frame bytes written at ingest and read by the encoder), fp32 in exact mode and on the host."""
    from ..runtime import compute_dtype
    return compute_dtype() if device_frames else torch.float32


class _PinnedRing:
    """256 pinned 256-byte slots allocated ONCE (a pinned allocation per prompt would bring its own stalls): slot i is reused
    only after the event recorded behind its previous copy has completed (normally long ago: 256 uploads earlier)."""
    SLOTS, SLOT_BYTES = 256, 256

    def __init__(self):
        self.buf = torch.empty(self.SLOTS * self.SLOT_BYTES, dtype=torch.uint8, pin_memory=True)
        self.events = [None] * self.SLOTS
        self.next = 0

    def stage(self, t, device):
        nbytes = t.numel() * t.element_size()
        if nbytes == 0 or nbytes > self.SLOT_BYTES:
            return None
        i, self.next = self.next, (self.next + 1) % self.SLOTS
        if self.events[i] is not None:
            self.events[i].synchronize()
        view = self.buf[i * self.SLOT_BYTES: i * self.SLOT_BYTES + nbytes].view(t.dtype).view(t.shape)
        view.copy_(t)
        out = view.to(device, non_blocking=True)
        if self.events[i] is None:
            self.events[i] = torch.cuda.Event()
        self.events[i].record(torch.cuda.current_stream(out.device))
        return out


_RINGS = {}                      # one ring per destination device (its events belong to that device)


def to_device_async(t, device):
    """Small host tensor (prompt coordinates, labels) -> device WITHOUT draining the stream: `cudaMemcpyAsync` from
    pageable memory synchronises the stream before the copy starts, which turns every prompt into a host<->GPU round
    trip (the reference does exactly that: `points.to(device)`, sam2_video_predictor.py:222-224).  Staged through a pinned
    slot the copy is a plain stream-ordered DMA; tensors beyond 256 bytes take the ordinary path."""
    dev = torch.device(device)
    if t.is_cuda or dev.type != "cuda" or os.environ.get("MS2_PINNED_PROMPTS", "1") == "0":
        return t.to(device)              # (MS2_PINNED_PROMPTS=0: the reference's pageable copy, for A/B measurements)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ring = _RINGS.get(key)
    if ring is None:
        ring = _RINGS[key] = _PinnedRing()
    out = ring.stage(t.contiguous(), dev)
    return out if out is not None else t.to(device)


class StreamedFrames:
    """Frames of a HOST tensor uploaded and normalised chunk by chunk on a side stream (the counterpart of the
    reference's AsyncVideoFrameLoader, utils/misc.py:92-160, for frames that are already decoded): indexing a frame
    makes the consumer's stream wait for the CUDA event of its chunk, so the host->device copy of later slices
    overlaps the encoding / tracking of earlier ones instead of preceding it.  `host`: [T,3,H,W] of any dtype
    (uint8 travels as uint8: a quarter of the fp32 bytes over PCIe) or uint8 [T,H,W,3] with nhwc=True."""

    def __init__(self, host, device, chunk=8, nhwc=False):
        assert not host.is_cuda and host.dim() == 4
        if (host.shape[-1] if nhwc else host.shape[1]) != 3:
            raise ValueError(f"StreamedFrames: expected {'[T,H,W,3]' if nhwc else '[T,3,H,W]'} frames, got {tuple(host.shape)}")
        if nhwc and host.dtype != torch.uint8:
            raise ValueError("StreamedFrames: channels-last frames must be uint8")
        self.T = host.shape[0]
        self.chunk = max(1, int(chunk))
        H, W = (host.shape[1], host.shape[2]) if nhwc else (host.shape[2], host.shape[3])
        self.out = torch.empty((self.T, 3, H, W), dtype=_frame_dtype(), device=device)
        self.stream = _upload_stream(device)
        self.out.record_stream(self.stream)
        self.stream.wait_stream(torch.cuda.current_stream(device))
        self.events, self.waited = [], []
        src = host if host.dtype in (torch.uint8, torch.float32) else host.float()
        with torch.cuda.stream(self.stream):
            for c0 in range(0, self.T, self.chunk):
                c1 = min(c0 + self.chunk, self.T)
                raw = src[c0:c1].to(device, non_blocking=True)
                ops.normalize_image(raw.contiguous(), out=self.out[c0:c1], nhwc=nhwc)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                self.events.append(ev)
                self.waited.append(False)

    def _wait(self, c):
        cur = torch.cuda.current_stream(self.out.device)
        if self.waited[c] is False:
            self.waited[c] = set()
        if cur.cuda_stream not in self.waited[c]:          # per consumer stream (tracking / slice-encoding streams)
            cur.wait_event(self.events[c])
            self.waited[c].add(cur.cuda_stream)

    def __len__(self):
        return self.T

    @property
    def shape(self):
        return self.out.shape

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            lo, hi, _ = idx.indices(self.T)
            for c in range(lo // self.chunk, (max(hi, lo + 1) - 1) // self.chunk + 1):
                self._wait(c)
        else:
            i = int(idx) % self.T
            self._wait(i // self.chunk)
        return self.out[idx]


class LazyHostFrames:
    """Frames that STAY on the host (`offload_video_to_cpu=True`) and are uploaded + normalised only when a frame is
    asked for (`async_loading_frames=True`): a rank that encodes one block of a long volume (parallel.
    encode_volume_sharded) moves only that block over PCIe.  `frames(f0, f1)` uploads a run of frames in one copy."""

    def __init__(self, host, device):
        assert not host.is_cuda and host.dim() == 4 and host.shape[1] == 3
        self.host = host if host.dtype in (torch.uint8, torch.float32) else host.float()
        self.device = torch.device(device)
        self.T = host.shape[0]

    def __len__(self):
        return self.T

    @property
    def shape(self):
        return self.host.shape

    def frames(self, f0, f1):
        raw = self.host[f0:f1].to(self.device, non_blocking=True)
        return ops.normalize_image(raw.contiguous(), nhwc=False, out_dtype=_frame_dtype())

    def __getitem__(self, idx):
        i = int(idx) % self.T
        return self.frames(i, i + 1)[0]


def load_video_frames_from_data(imgs_tensor, offload_video_to_cpu=False, img_mean=(0.485, 0.456, 0.406),
                                img_std=(0.229, 0.224, 0.225), async_loading_frames=False, device="cuda"):
    """utils/misc.py:215-244: [T,3,S,S] in 0..255 (any dtype, like the reference's `imgs_tensor / 255.0`) -> normalised
    fp32 frames, one fused kernel on device.  `async_loading_frames=True` with a host tensor streams the upload
    (StreamedFrames)."""
    assert tuple(img_mean) == (0.485, 0.456, 0.406) and tuple(img_std) == (0.229, 0.224, 0.225)
    x = imgs_tensor
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"load_video_frames_from_data: expected imgs_tensor [T,3,H,W], got {tuple(x.shape)}")
    if async_loading_frames and not x.is_cuda and torch.cuda.is_available():
        return LazyHostFrames(x, device) if offload_video_to_cpu else StreamedFrames(x, device)
    if not x.is_cuda:
        x = x.to(device, non_blocking=True)
    if x.dtype == torch.uint8:
        images = ops.normalize_image(x.contiguous(), nhwc=False)
    else:
        images = ops.normalize_image(x.float().contiguous())
    return images.cpu() if offload_video_to_cpu else images


def _decode_jpeg(path, image_size):
    """utils/misc.py:92-101 `_load_img_as_tensor` up to the uint8 array: PIL decode -> RGB -> PIL resize (default
    resampling).  PIL releases the GIL while decoding / resizing, so a thread pool scales over host cores."""
    from PIL import Image
    pil = Image.open(path)
    arr = np.asarray(pil.convert("RGB").resize((image_size, image_size)))
    if arr.dtype != np.uint8:
        raise RuntimeError(f"Unknown image dtype: {arr.dtype} on {path}")
    w, h = pil.size
    return arr, h, w


def _jpeg_paths(video_path):
    if not (isinstance(video_path, str) and os.path.isdir(video_path)):
        raise NotImplementedError("Only JPEG frames are supported at this moment")
    names = [p for p in os.listdir(video_path) if os.path.splitext(p)[-1] in (".jpg", ".jpeg", ".JPG", ".JPEG")]
    names.sort(key=lambda p: int(os.path.splitext(p)[0]))
    if not names:
        raise RuntimeError(f"no images found in {video_path}")
    return [os.path.join(video_path, n) for n in names]


class AsyncVideoFrameLoader:
    """utils/misc.py:104-160 re-designed for a GPU consumer.  The reference decodes every frame on ONE background thread
    into fp32 tensors and uploads them one by one from that thread.  Here a pool of host threads decodes "<index>.jpg"
    files (PIL, GIL released) straight into ONE pinned uint8 [T,S,S,3] staging buffer; finished chunks of `chunk`
    frames are uploaded as uint8 (a quarter of the fp32 bytes) on the upload stream and normalised there by
    `ms2_normalize_image` (bf16 frames in bf16 mode).  All CUDA calls stay on the caller's thread (the decode threads
    never touch CUDA, so they cannot disturb a CUDA-graph capture): `loader[i]` enqueues the uploads of every chunk that
    finished decoding, blocks on the host only if chunk(i) is still being decoded, and makes the consumer's stream wait
    for that chunk's CUDA event.  Frame 0 is decoded synchronously (it fills video_height / video_width, as in the
    reference).  With `offload_video_to_cpu=True` frames stay on the host as fp32 (the reference's layout)."""

    def __init__(self, img_paths, image_size, offload_video_to_cpu=False, device="cuda", chunk=8, workers=None):
        import threading
        from concurrent.futures import ThreadPoolExecutor
        self.img_paths = list(img_paths)
        self.image_size = image_size
        self.offload_video_to_cpu = offload_video_to_cpu
        self.device = torch.device(device)
        self.T = len(self.img_paths)
        self.chunk = max(1, int(chunk))
        self.exception = None
        on_gpu = (not offload_video_to_cpu) and self.device.type == "cuda"
        self.staging = torch.empty((self.T, image_size, image_size, 3), dtype=torch.uint8)
        if on_gpu:
            self.staging = self.staging.pin_memory()
            self.out = torch.empty((self.T, 3, image_size, image_size), dtype=_frame_dtype(), device=self.device)
            self.stream = _upload_stream(self.device)
            self.out.record_stream(self.stream)
        else:
            self.out = torch.empty((self.T, 3, image_size, image_size), dtype=torch.float32)
        self._np = self.staging.numpy()
        n_chunks = (self.T + self.chunk - 1) // self.chunk
        self._left = [min(self.chunk, self.T - c * self.chunk) for c in range(n_chunks)]     # frames still decoding
        self._decoded = [threading.Event() for _ in range(n_chunks)]
        self._uploaded = [False] * n_chunks
        self._events = [None] * n_chunks
        self._waited = [set() for _ in range(n_chunks)]
        self._lock = threading.Lock()
        self._one(0)                                          # synchronously: video size + the frame users click first
        workers = workers or min(8, os.cpu_count() or 1)
        self._pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="ms2-jpeg")
        for i in range(1, self.T):
            self._pool.submit(self._one, i)
        self._pool.shutdown(wait=False)

    def _one(self, i):
        try:
            arr, h, w = _decode_jpeg(self.img_paths[i], self.image_size)
            self._np[i] = arr
            if i == 0:
                self.video_height, self.video_width = h, w
        except Exception as e:                                 # surfaces in the consumer (utils/misc.py:128-134)
            self.exception = e
        finally:
            c = i // self.chunk
            with self._lock:
                self._left[c] -= 1
                if self._left[c] == 0:
                    self._decoded[c].set()

    def _upload(self, c):
        c0, c1 = c * self.chunk, min((c + 1) * self.chunk, self.T)
        if self.offload_video_to_cpu or self.device.type != "cuda":
            x = self.staging[c0:c1].permute(0, 3, 1, 2).float() / 255.0
            mean = torch.tensor((0.485, 0.456, 0.406))[:, None, None]
            std = torch.tensor((0.229, 0.224, 0.225))[:, None, None]
            self.out[c0:c1] = (x - mean) / std
        else:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                raw = self.staging[c0:c1].to(self.device, non_blocking=True)
                ops.normalize_image(raw, out=self.out[c0:c1], nhwc=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self._events[c] = ev
        self._uploaded[c] = True

    def _ready(self, c):
        if self.exception is not None:
            raise RuntimeError("Failure in frame loading thread") from self.exception
        for k in range(len(self._uploaded)):                   # push every finished chunk to the GPU, in order, non-blocking
            if not self._uploaded[k] and (k == c or self._decoded[k].is_set()):
                if k == c:
                    self._decoded[k].wait()
                    if self.exception is not None:
                        raise RuntimeError("Failure in frame loading thread") from self.exception
                self._upload(k)
        if self._events[c] is not None:
            cur = torch.cuda.current_stream(self.device)
            if cur.cuda_stream not in self._waited[c]:
                cur.wait_event(self._events[c])
                self._waited[c].add(cur.cuda_stream)

    def __getitem__(self, index):
        i = int(index) % self.T
        self._ready(i // self.chunk)
        return self.out[i]

    def __len__(self):
        return self.T

    @property
    def shape(self):
        return self.out.shape


def load_video_frames(video_path, image_size, offload_video_to_cpu=False, img_mean=(0.485, 0.456, 0.406),
                      img_std=(0.229, 0.224, 0.225), async_loading_frames=False, device="cuda"):
    """utils/misc.py:163-212: a directory of "<frame_index>.jpg" files -> (frames, video_height, video_width).
    async_loading_frames=True returns the lazy `AsyncVideoFrameLoader`; otherwise all frames are decoded (thread pool),
    uploaded as uint8 and normalised by one kernel launch."""
    assert tuple(img_mean) == (0.485, 0.456, 0.406) and tuple(img_std) == (0.229, 0.224, 0.225)
    paths = _jpeg_paths(video_path)
    lazy = AsyncVideoFrameLoader(paths, image_size, offload_video_to_cpu, device=device,
                                 chunk=8 if async_loading_frames else len(paths))
    if async_loading_frames:
        return lazy, lazy.video_height, lazy.video_width
    lazy[len(paths) - 1]                                       # decode + upload everything now
    return lazy.out, lazy.video_height, lazy.video_width


def concat_points(old_point_inputs, new_points, new_labels):
    """utils/misc.py:261-269."""
    if old_point_inputs is None:
        points, labels = new_points, new_labels
    else:
        points = torch.cat([old_point_inputs["point_coords"], new_points], dim=1)
        labels = torch.cat([old_point_inputs["point_labels"], new_labels], dim=1)
    return {"point_coords": points, "point_labels": labels}
