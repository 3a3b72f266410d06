"""Small shared pieces: the MLP / LayerNorm parameter containers (state_dict-compatible with the
reference's modeling/sam2_utils.py:107-147), layout helpers between the reference-facing NCHW /
sequence-first tensors and the token-major tensors the kernels use, and frame-selection helpers."""
import torch
from torch import nn

from .. import ops
from ..runtime import compute_dtype, w_c, p32


class Linear(nn.Linear):
    """nn.Linear parameters; the contraction runs in ms2_gemm."""

    def forward(self, x, out_dtype=torch.float32, act=ops.ACT_NONE, residual=None, colscale=None, out=None):
        return ops.gemm(to_compute(x), w_c(self.weight), p32(self.bias), out_dtype=out_dtype, act=act,
                        residual=residual, colscale=colscale, out=out)


class LayerNorm(nn.LayerNorm):
    def forward(self, x, out_dtype=torch.float32, add=None, act=ops.ACT_NONE):
        return ops.layernorm(x, p32(self.weight), p32(self.bias), self.eps, out_dtype=out_dtype, add=add, act=act)


class LayerNorm2d(nn.Module):
    """Channel LayerNorm of NCHW maps in the reference (sam2_utils.py:133-147); token-major here."""

    def __init__(self, num_channels, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps

    def forward(self, x_nhwc, out_dtype=torch.float32, act=ops.ACT_NONE):
        return ops.layernorm(x_nhwc, p32(self.weight), p32(self.bias), self.eps, out_dtype=out_dtype, act=act)


class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers, activation=None, sigmoid_output=False):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
        self.sigmoid_output = sigmoid_output
        self.act = ops.ACT_RELU if activation in (None, "relu") else ops.ACT_GELU

    def forward(self, x, residual=None):
        cd = compute_dtype()
        for i, layer in enumerate(self.layers):
            last = i == self.num_layers - 1
            if last:
                x = layer(x, out_dtype=torch.float32, act=ops.ACT_SIGMOID if self.sigmoid_output else ops.ACT_NONE,
                          residual=residual)
            else:
                x = layer(x, out_dtype=cd, act=self.act)
        return x


def to_compute(x):
    return ops.cast(x, compute_dtype()) if x.dtype != compute_dtype() else x


def as_nhwc(x):
    """NCHW-shaped fp32 tensor -> contiguous NHWC tensor (free when x is channels-last strided)."""
    if x.dtype != torch.float32:
        x = x.float()
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return v
    if x.is_contiguous():
        return ops.nchw_to_nhwc(x)
    return v.contiguous()          # expanded / oddly-strided API input: one plain copy


def as_nchw_view(x_nhwc):
    """contiguous NHWC -> NCHW-shaped (channels-last strided) view, no copy."""
    return x_nhwc.permute(0, 3, 1, 2)


def seq_to_tokens(x):
    """[L,B,C] sequence-first -> contiguous [B,L,C]."""
    v = x.permute(1, 0, 2)
    return v if v.is_contiguous() else v.contiguous()


def select_closest_cond_frames(frame_idx, cond_frame_outputs, max_cond_frame_num):
    """Conditioning-frame selection rule of the reference (modeling/sam2_utils.py:15-57)."""
    if max_cond_frame_num == -1 or len(cond_frame_outputs) <= max_cond_frame_num:
        return cond_frame_outputs, {}
    assert max_cond_frame_num >= 2, "we should allow using 2+ conditioning frames"
    picked = {}
    before = [t for t in cond_frame_outputs if t < frame_idx]
    after = [t for t in cond_frame_outputs if t >= frame_idx]
    if before:
        picked[max(before)] = cond_frame_outputs[max(before)]
    if after:
        picked[min(after)] = cond_frame_outputs[min(after)]
    rest = sorted((t for t in cond_frame_outputs if t not in picked), key=lambda t: abs(t - frame_idx))
    for t in rest[: max_cond_frame_num - len(picked)]:
        picked[t] = cond_frame_outputs[t]
    return picked, {t: v for t, v in cond_frame_outputs.items() if t not in picked}


def get_1d_sine_pe(pos_inds, dim, temperature=10000):
    """1-D sine embedding (modeling/sam2_utils.py:60-71); tiny host-side table math."""
    pe_dim = dim // 2
    dim_t = torch.arange(pe_dim, dtype=torch.float32, device=pos_inds.device)
    dim_t = temperature ** (2 * (dim_t // 2) / pe_dim)
    pos_embed = pos_inds.unsqueeze(-1) / dim_t
    return torch.cat([pos_embed.sin(), pos_embed.cos()], dim=-1)
