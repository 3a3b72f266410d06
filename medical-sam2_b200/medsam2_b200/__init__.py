"""medsam2_b200 — B200-native (sm_100a) implementation of Medical-SAM2's per-slice inference hot
path behind the reference's Python API (build_sam2 / SAM2ImagePredictor / SAM2VideoPredictor and
the `_C.get_connected_componnets` native boundary)."""
from .build_sam import build_sam2, build_sam2_video_predictor  # noqa: F401
from .runtime import compute, compute_dtype, set_compute_dtype  # noqa: F401
from .sam2_image_predictor import SAM2ImagePredictor  # noqa: F401
from .sam2_video_predictor import SAM2VideoPredictor  # noqa: F401

__version__ = "0.1.0"
