"""Python wrappers of the C-ABI ops (include/mgld.h).  Tensors are CUDA fp16 NHWC unless stated otherwise.

These wrappers only marshal pointers/sizes into the C structs; all arithmetic happens in libmgld.so.
"""
import ctypes

import torch

from . import lib as _L

TAPS_1, TAPS_T3, TAPS_3X3 = 1, 3, 9
EPI_LINEAR, EPI_GEGLU, EPI_SPADE = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SILU, ACT_LRELU02, ACT_GELU = 0, 1, 2, 3, 4


# ---------------------------------------------------------------------------------------------------------------
# one-time weight layout conversion (SURVEY.md §8b "weights_pack_*")
# ---------------------------------------------------------------------------------------------------------------
def pack_conv_weight(w):
    """[Cout, Cin, kh, kw] (torch conv2d) -> [Cout, kh*kw*Cin] fp16, tap-major / channel-fastest."""
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).to(torch.float16).contiguous()


def pack_temporal_weight(w):
    """[C, C, 3, 1, 1] (torch conv3d, util.py:296) -> [C, 3*C] fp16, time-tap-major."""
    co, ci = w.shape[:2]
    return w[:, :, :, 0, 0].permute(0, 2, 1).reshape(co, 3 * ci).to(torch.float16).contiguous()


def interleave_pair(wa, wb, blk=64):
    """Interleave the rows of two [N, ...] tensors in blocks of 64: [a0..a63 | b0..b63 | a64.. ]."""
    n = wa.shape[0]
    assert wb.shape == wa.shape and n % blk == 0
    rest = wa.shape[1:]
    return torch.stack([wa.reshape(n // blk, blk, *rest), wb.reshape(n // blk, blk, *rest)], dim=1) \
        .reshape(2 * n, *rest).contiguous()


# ---------------------------------------------------------------------------------------------------------------
def conv_gemm(a, w, *, taps=TAPS_1, a2=None, bias=None, epilogue=EPI_LINEAR, act=ACT_NONE, alpha=1.0, beta=0.0,
              res=None, h=None, gn_stats=None, gn_weight=None, gn_bias=None, groups=32, out=None, out_col0=0,
              out_f32=False, block_n=0):
    """Implicit-GEMM conv / linear on tcgen05 (mgld_conv_gemm).

    a: [T,H,W,C1] or [M,C1] fp16 (last dim contiguous);  w: packed [N, taps*(C1+C2)] fp16.
    Returns out [T,H,W,n_out] (or [M,n_out]).
    """
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and a.is_cuda
    if a.dim() == 2:
        T, H, W = 1, 1, a.shape[0]
    else:
        T, H, W = a.shape[0], a.shape[1], a.shape[2]
    C1 = a.shape[-1]
    assert a.stride(-1) == 1
    lda = a.stride(-2)
    if a.dim() == 4:
        assert a.stride(1) == lda * W and a.stride(0) == lda * W * H
    C2, lda2 = 0, 0
    if a2 is not None:
        assert a2.shape[:-1] == a.shape[:-1] and a2.stride(-1) == 1
        C2, lda2 = a2.shape[-1], a2.stride(-2)
    N = w.shape[0]
    assert w.shape[1] == taps * (C1 + C2), (w.shape, taps, C1, C2)
    assert w.is_contiguous()
    n_out = N // 2 if epilogue in (EPI_GEGLU, EPI_SPADE) else N
    M = T * H * W
    if out is None:
        out = torch.empty(*a.shape[:-1], n_out, device=a.device, dtype=torch.float32 if out_f32 else torch.float16)
        out_col0 = 0
    ldout = out.stride(-2)
    d = _L.ConvGemmDesc()
    d.a, d.a2 = a.data_ptr(), (a2.data_ptr() if a2 is not None else None)
    d.T, d.H, d.W, d.C1, d.C2, d.lda, d.lda2 = T, H, W, C1, C2, lda, lda2
    d.w, d.N, d.taps, d.block_n = w.data_ptr(), N, taps, block_n
    d.epilogue, d.act = epilogue, act
    d.bias = bias.data_ptr() if bias is not None else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    d.alpha, d.beta = alpha, beta
    d.res = res.data_ptr() if res is not None else None
    d.ldres = res.stride(-2) if res is not None else 0
    d.h = h.data_ptr() if h is not None else None
    d.ldh = h.stride(-2) if h is not None else 0
    d.gn_stats = gn_stats.data_ptr() if gn_stats is not None else None
    d.gn_weight = gn_weight.data_ptr() if gn_weight is not None else None
    d.gn_bias = gn_bias.data_ptr() if gn_bias is not None else None
    d.groups = groups
    d.out, d.ldout, d.out_col0, d.out_f32 = out.data_ptr(), ldout, out_col0, int(out_f32)
    _L.check(_L.lib().mgld_conv_gemm(ctypes.byref(d), _L.stream_ptr()))
    return out


# ---------------------------------------------------------------------------------------------------------------
def attention(q, k, v, *, batch, heads, head_dim, nq, nkv, scale, q_col0=0, k_col0=0, v_col0=0,
              q_head_stride=None, k_head_stride=None, v_head_stride=None, kv_batched=True, out=None):
    """Fused softmax attention (mgld_attention).  q/k/v: fp16 2-D matrices [batch*n, ld] (may be the same buffer);
    head h lives at columns [col0 + h*head_stride, +head_dim).  Returns out [batch*nq, heads*head_dim] fp16."""
    for t in (q, k, v):
        assert t.dtype == torch.float16 and t.is_cuda and t.dim() == 2 and t.stride(1) == 1
    assert q.shape[0] >= batch * nq and k.shape[0] >= (batch if kv_batched else 1) * nkv
    if out is None:
        out = torch.empty(batch * nq, heads * head_dim, device=q.device, dtype=torch.float16)
    d = _L.AttentionDesc()
    d.q, d.k, d.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    d.ldq, d.ldk, d.ldv = q.stride(0), k.stride(0), v.stride(0)
    d.q_col0, d.k_col0, d.v_col0 = q_col0, k_col0, v_col0
    d.q_head_stride = head_dim if q_head_stride is None else q_head_stride
    d.k_head_stride = head_dim if k_head_stride is None else k_head_stride
    d.v_head_stride = head_dim if v_head_stride is None else v_head_stride
    d.batch, d.heads, d.head_dim, d.nq, d.nkv = batch, heads, head_dim, nq, nkv
    d.kv_batched, d.scale = int(kv_batched), float(scale)
    d.out, d.ldo = out.data_ptr(), out.stride(0)
    _L.check(_L.lib().mgld_attention(ctypes.byref(d), _L.stream_ptr()))
    return out


# ---------------------------------------------------------------------------------------------------------------
# flow ops (fp32 NCHW)
# ---------------------------------------------------------------------------------------------------------------
def _f32c(t):
    assert t.is_cuda and t.dtype == torch.float32
    return t.contiguous()


def flow_warp_f32(x, flow, flow_layout=0, nearest=False, border=False, align_corners=True):
    x, flow = _f32c(x), _f32c(flow)
    n, c, h, w = x.shape
    out = torch.empty_like(x)
    _L.check(_L.lib().mgld_flow_warp_f32(_L.ptr(x), _L.ptr(flow), _L.ptr(out), n, c, h, w, int(flow_layout),
                                         int(nearest), int(border), int(align_corners), _L.stream_ptr()))
    return out


def flow_warp_bwd_input_f32(grad_out, flow, flow_layout=0, border=False, align_corners=True):
    grad_out, flow = _f32c(grad_out), _f32c(flow)
    n, c, h, w = grad_out.shape
    gin = torch.empty_like(grad_out)
    _L.check(_L.lib().mgld_flow_warp_bwd_input_f32(_L.ptr(grad_out), _L.ptr(flow), _L.ptr(gin), n, c, h, w,
                                                   int(flow_layout), int(border), int(align_corners),
                                                   _L.stream_ptr()))
    return gin


def fb_consistency_f32(fwd_flow, bwd_flow, alpha=0.01, beta=0.5):
    fwd_flow, bwd_flow = _f32c(fwd_flow), _f32c(bwd_flow)
    b, _, h, w = fwd_flow.shape
    fo = torch.empty(b, h, w, device=fwd_flow.device, dtype=torch.float32)
    bo = torch.empty_like(fo)
    _L.check(_L.lib().mgld_fb_consistency_f32(_L.ptr(fwd_flow), _L.ptr(bwd_flow), _L.ptr(fo), _L.ptr(bo), b, h, w,
                                              ctypes.c_float(alpha), ctypes.c_float(beta), _L.stream_ptr()))
    return fo, bo


def motion_guidance_f32(latents, flow_fwd_prop, flow_bwd_prop, fwd_occ, bwd_occ, step, want_loss=False):
    """latents (t,c,h,w); flows (t-1,2,h,w); occs (t-1,h,w).  Returns latents - step * grad (and the loss)."""
    latents = _f32c(latents)
    t, c, h, w = latents.shape
    ws = torch.empty_like(latents)
    out = torch.empty_like(latents)
    loss = torch.empty(1, device=latents.device, dtype=torch.float32) if want_loss else None
    args = [_f32c(a) if a is not None else None for a in (flow_fwd_prop, flow_bwd_prop, fwd_occ, bwd_occ)]
    _L.check(_L.lib().mgld_motion_guidance_f32(_L.ptr(latents), _L.ptr(args[0]), _L.ptr(args[1]), _L.ptr(args[2]),
                                               _L.ptr(args[3]), _L.ptr(ws), _L.ptr(out), _L.ptr(loss),
                                               ctypes.c_float(step), t, c, h, w, _L.stream_ptr()))
    return (out, loss, ws) if want_loss else out


def resize_flow_f32(flow, oh, ow):
    flow = _f32c(flow)
    n, _, h, w = flow.shape
    out = torch.empty(n, 2, oh, ow, device=flow.device, dtype=torch.float32)
    _L.check(_L.lib().mgld_resize_flow_f32(_L.ptr(flow), _L.ptr(out), n, h, w, oh, ow, _L.stream_ptr()))
    return out
