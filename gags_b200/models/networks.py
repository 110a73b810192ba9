"""CNN_decoder / CNN_scale_decoder of /root/reference/models/networks.py:109-248 — the two 1x1-conv
stacks train.py applies to the rendered feature map (train.py:149 scale decoder, :159 feature
decoder) — evaluated on the CHANNEL-LAST raster the rasteriser produces.

A 1x1 convolution is a per-pixel linear map.  render()["render"] is a [D,H,W] *view* of the
channel-last [H,W,D] buffer, so the whole stack is a chain of plain GEMMs on the [H*W, D] matrix
that already sits in HBM: no NCHW copy of the 2 GB raster (what cuDNN would need), bias + ReLU in
place, the residual adds of the reference kept, and the final L2-normalise / softmax over channels
done on the contiguous last dimension.  The GEMMs go to cuBLAS through torch (plain library GEMMs);
fp32 accumulate, TF32 inputs only if the caller enabled them globally, as for the reference's convs.

Parameter layout is the reference's (`decoder.{i}.weight` [out,in,1,1], `decoder.{i}.bias`), so
`{module_state_dict, optimizer_state_dict}` checkpoints (train.py:235-242) load unchanged.
Construction takes `device` instead of hard-coding .cuda().
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


def _as_rows(x: torch.Tensor):
    """[C,H,W] (any strides) -> ([H*W, C] matrix, H, W) without a copy when x is a permuted view of a
    channel-last buffer."""
    if x.dim() != 3:
        raise ValueError("expected a [C,H,W] feature map")
    c, h, w = x.shape
    rows = x.permute(1, 2, 0)
    rows = rows.reshape(h * w, c) if rows.is_contiguous() else rows.contiguous().reshape(h * w, c)
    return rows, h, w


def _lin(conv: nn.Conv2d, rows: torch.Tensor, relu: bool) -> torch.Tensor:
    out = torch.addmm(conv.bias, rows, conv.weight.view(conv.out_channels, conv.in_channels).t())
    return F.relu_(out) if relu else out


def _stack(dims_in, dims_out, device):
    layers = []
    for i, (a, b) in enumerate(zip(dims_in, dims_out)):
        if i > 0:
            layers.append(nn.ReLU())
        layers.append(nn.Conv2d(a, b, kernel_size=1))
    return nn.ModuleList(layers).to(device)


class CNN_decoder(nn.Module):
    """16 -> 256 x 8 -> output_dim with two residual adds and a channel L2-normalise (:109-218)."""

    def __init__(self, input_dim, output_dim, device="cuda"):
        super().__init__()
        self.decoder = _stack([input_dim] + [256] * 8, [256] * 8 + [output_dim], device)

    def forward(self, x):
        d = self.decoder
        rows, h, w = _as_rows(x)
        x1 = _lin(d[0], rows, True)
        x2 = _lin(d[4], _lin(d[2], x1, True), True)
        x3 = _lin(d[6], x1 + x2, True)
        x4 = _lin(d[10], _lin(d[8], x3, True), True)
        x5 = _lin(d[16], _lin(d[14], _lin(d[12], x3 + x4, True), True), False)
        out = F.normalize(x5, dim=-1)                     # == F.normalize(., dim=0) of [C,H,W]
        return out.view(h, w, -1).permute(2, 0, 1)       # [C,H,W] view of the channel-last result


class CNN_scale_decoder(nn.Module):
    """input_dim -> 64 -> 128 -> 64 -> 32 -> 16 -> output_dim, softmax over channels (:220-248)."""

    def __init__(self, input_dim, output_dim, device="cuda"):
        super().__init__()
        hidden = [64, 128, 64, 32, 16, output_dim]
        self.decoder = _stack([input_dim] + hidden[:-1], hidden, device)

    def forward(self, x):
        rows, h, w = _as_rows(x)
        convs = [m for m in self.decoder if isinstance(m, nn.Conv2d)]
        for i, c in enumerate(convs):
            rows = _lin(c, rows, i + 1 < len(convs))
        out = F.softmax(rows, dim=-1)
        return out.view(h, w, -1).permute(2, 0, 1)
