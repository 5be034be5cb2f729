"""Flow utilities with the reference's signatures, running on the mgld kernels (fp32 NCHW tensors on CUDA).

  flow_warp                           basicsr/archs/arch_util.py:156-194
  resize_flow                         basicsr/archs/arch_util.py:235-270
  forward_backward_consistency_check  scripts/util_flow.py:114-136
"""
import torch

from . import ops as _ops


def flow_warp(x, flow, interp_mode="bilinear", padding_mode="zeros", align_corners=True, return_mask=False, ops=None):
    """Warp (n,c,h,w) `x` with a pixel-unit flow (n,h,w,2)."""
    ops = ops or _ops
    assert x.size()[-2:] == flow.size()[1:3]                      # arch_util.py:172
    if interp_mode not in ("bilinear", "nearest"):
        raise ValueError(f"interp_mode {interp_mode!r} not supported")
    if padding_mode not in ("zeros", "border"):
        raise NotImplementedError("padding_mode 'reflection' is not used by the hot path")
    out = ops.flow_warp_f32(x.float(), flow.float(), 0, nearest=interp_mode == "nearest",
                            border=padding_mode == "border", align_corners=align_corners)
    if not return_mask:
        return out
    mask = ops.flow_warp_f32(torch.ones_like(x, dtype=torch.float32), flow.float(), 0,
                             nearest=interp_mode == "nearest", border=padding_mode == "border",
                             align_corners=align_corners)
    mask = (mask >= 0.9999).float()                               # arch_util.py:191-193
    return out, mask


def resize_flow(flow, size_type, sizes, interp_mode="bilinear", align_corners=False, ops=None):
    ops = ops or _ops
    _, _, flow_h, flow_w = flow.size()
    if size_type == "ratio":
        output_h, output_w = int(flow_h * sizes[0]), int(flow_w * sizes[1])
    elif size_type == "shape":
        output_h, output_w = sizes[0], sizes[1]
    else:
        raise ValueError(f"Size type should be ratio or shape, but got type {size_type}.")
    if interp_mode != "bilinear" or align_corners:
        raise NotImplementedError("only the reference's default (bilinear, align_corners=False) is implemented")
    return ops.resize_flow_f32(flow.float(), output_h, output_w)


def forward_backward_consistency_check(fwd_flow, bwd_flow, alpha=0.01, beta=0.5, ops=None):
    ops = ops or _ops
    assert fwd_flow.dim() == 4 and bwd_flow.dim() == 4            # util_flow.py:121-122
    assert fwd_flow.size(1) == 2 and bwd_flow.size(1) == 2
    return ops.fb_consistency_f32(fwd_flow.float(), bwd_flow.float(), alpha, beta)
