"""Drop-in for the reference's pybind module `sam2_train._C` (csrc/connected_components.cu:284-289).

`get_connected_componnets` (spelling preserved) keeps the reference contract — CUDA uint8
[N,1,H,W] in, `[labels, counts]` int32 out, even H/W, current stream, errors as RuntimeError — and
runs the single-launch sm_100a kernel of csrc/cc.cu through the C-ABI.
"""
import torch

from . import ops
from .native import NativeError


def get_connected_componnets(inputs):
    if not isinstance(inputs, torch.Tensor) or not inputs.is_cuda:
        raise RuntimeError("inputs must be a CUDA tensor")
    if inputs.dim() != 4:
        raise RuntimeError("inputs must be [N, 1, H, W] shape")
    if inputs.dtype != torch.uint8:
        raise RuntimeError("inputs must be a uint8 type")
    if inputs.shape[1] != 1:
        raise RuntimeError("inputs must be [N, 1, H, W] shape")
    try:
        labels, counts = ops.cc_label(inputs.contiguous())
    except NativeError as e:
        raise RuntimeError(str(e)) from e
    return [labels, counts]


get_connected_components = get_connected_componnets
