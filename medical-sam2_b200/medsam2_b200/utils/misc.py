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


class StreamedFrames:
    """Frames of a HOST tensor uploaded and normalised chunk by chunk on a side stream (the counterpart of the
    reference's AsyncVideoFrameLoader, utils/misc.py:92-160, for frames that are already decoded): indexing a frame
    makes the consumer's stream wait for the CUDA event of its chunk, so the host->device copy of later slices
    overlaps the encoding / tracking of earlier ones instead of preceding it."""

    def __init__(self, host, device, chunk=8):
        assert not host.is_cuda and host.dim() == 4
        self.T = host.shape[0]
        self.chunk = max(1, int(chunk))
        layout_u8 = host.dtype == torch.uint8
        H, W = (host.shape[1], host.shape[2]) if layout_u8 else (host.shape[2], host.shape[3])
        self.out = torch.empty((self.T, 3, H, W), dtype=torch.float32, device=device)
        self.stream = _upload_stream(device)
        self.out.record_stream(self.stream)
        self.stream.wait_stream(torch.cuda.current_stream(device))
        self.events, self.waited = [], []
        src = host if (layout_u8 or host.dtype == torch.float32) else host.float()
        with torch.cuda.stream(self.stream):
            for c0 in range(0, self.T, self.chunk):
                c1 = min(c0 + self.chunk, self.T)
                raw = src[c0:c1].to(device, non_blocking=True)
                ops.normalize_image(raw.contiguous(), out=self.out[c0:c1])
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


def load_video_frames_from_data(imgs_tensor, offload_video_to_cpu=False, img_mean=(0.485, 0.456, 0.406),
                                img_std=(0.229, 0.224, 0.225), async_loading_frames=False, device="cuda"):
    """utils/misc.py:215-244: [T,3,S,S] in 0..255 -> normalised fp32 frames, one fused kernel on device.
    `async_loading_frames=True` with a host tensor streams the upload (StreamedFrames)."""
    assert tuple(img_mean) == (0.485, 0.456, 0.406) and tuple(img_std) == (0.229, 0.224, 0.225)
    x = imgs_tensor
    if async_loading_frames and not x.is_cuda and not offload_video_to_cpu and torch.cuda.is_available():
        return StreamedFrames(x, device)
    if not x.is_cuda:
        x = x.to(device, non_blocking=True)
    images = ops.normalize_image(x.float().contiguous() if x.dtype != torch.uint8 else x.contiguous())
    return images.cpu() if offload_video_to_cpu else images


def load_video_frames(video_path, image_size, offload_video_to_cpu=False, img_mean=(0.485, 0.456, 0.406),
                      img_std=(0.229, 0.224, 0.225), async_loading_frames=False, device="cuda"):
    """utils/misc.py:163-212: a directory of <frame_index>.jpg files -> (frames, H, W)."""
    from PIL import Image
    if not (isinstance(video_path, str) and os.path.isdir(video_path)):
        raise NotImplementedError("Only JPEG frames are supported at this moment")
    names = [p for p in os.listdir(video_path) if os.path.splitext(p)[-1] in (".jpg", ".jpeg", ".JPG", ".JPEG")]
    names.sort(key=lambda p: int(os.path.splitext(p)[0]))
    if not names:
        raise RuntimeError(f"no images found in {video_path}")
    frames = []
    for n in names:
        img = Image.open(os.path.join(video_path, n))
        W0, H0 = img.size
        frames.append(np.array(img.convert("RGB").resize((image_size, image_size))))
    x = torch.from_numpy(np.stack(frames)).to(device)                      # uint8 [T,S,S,3]
    images = ops.normalize_image(x)
    return (images.cpu() if offload_video_to_cpu else images), H0, W0


def concat_points(old_point_inputs, new_points, new_labels):
    """utils/misc.py:261-269."""
    if old_point_inputs is None:
        points, labels = new_points, new_labels
    else:
        points = torch.cat([old_point_inputs["point_coords"], new_points], dim=1)
        labels = torch.cat([old_point_inputs["point_labels"], new_labels], dim=1)
    return {"point_coords": points, "point_labels": labels}
