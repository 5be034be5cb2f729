"""TEST INFRASTRUCTURE: torch (CPU-capable) emulation of every op in mgld_vsr_b200.ops, same signatures.

Lets the CPU test-suite exercise the *host graph* (weight packing, layouts, layer order, fused-epilogue usage) against
the oracle without a GPU.  Math in fp32 on the fp16-stored operands, outputs rounded to the op's output dtype —
i.e. the numerics contract of the CUDA kernels.  Never imported by the product package.
"""
import math

import torch
import torch.nn.functional as F

LAUNCHES = [0]
TAPS_1, TAPS_T3, TAPS_3X3, TAPS_1X5, TAPS_5X1 = 1, 3, 9, 5, 6
EPI_LINEAR, EPI_GEGLU, EPI_SPADE = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SILU, ACT_LRELU02, ACT_GELU, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4, 5, 6

from mgld_vsr_b200.ops import pack_conv_weight, pack_temporal_weight, interleave_pair  # pure-torch layout helpers


def _act(v, act):
    if act == ACT_RELU: return F.relu(v)
    if act == ACT_SILU: return F.silu(v)
    if act == ACT_LRELU02: return F.leaky_relu(v, 0.2)
    if act == ACT_GELU: return F.gelu(v)
    if act == ACT_SIGMOID: return torch.sigmoid(v)
    if act == ACT_TANH: return torch.tanh(v)
    return v


def _deinterleave(t, blk=64):
    n = t.shape[0] // 2
    r = t.reshape(n // blk, 2, blk, *t.shape[1:])
    return r[:, 0].reshape(n, *t.shape[1:]), r[:, 1].reshape(n, *t.shape[1:])


def conv_gemm(a, w, *, taps=TAPS_1, a2=None, bias=None, epilogue=EPI_LINEAR, act=ACT_NONE, alpha=1.0, beta=0.0,
              res=None, h=None, gn_stats=None, gn_weight=None, gn_bias=None, groups=32, out=None, out_col0=0,
              out_f32=False, block_n=0, stats_out=None, stats_groups=32):
    x = a if a2 is None else torch.cat([a, a2], dim=-1)
    shp = x.shape[:-1]
    C = x.shape[-1]
    xf = x.float()
    wf = w.float()
    N = w.shape[0]
    if taps == 1:
        acc = xf.reshape(-1, C) @ wf.t()
    elif taps in (TAPS_1X5, TAPS_5X1):
        T, H, W = shp
        kh, kw = (1, 5) if taps == TAPS_1X5 else (5, 1)
        wt = wf.reshape(N, kh, kw, C).permute(0, 3, 1, 2)
        acc = F.conv2d(xf.permute(0, 3, 1, 2), wt, padding=(kh // 2, kw // 2)).permute(0, 2, 3, 1).reshape(-1, N)
    elif taps == 9:
        T, H, W = shp
        wt = wf.reshape(N, 3, 3, C).permute(0, 3, 1, 2)
        acc = F.conv2d(xf.permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    else:
        T, H, W = shp
        wt = wf.reshape(N, 3, C).permute(0, 2, 1)[:, :, :, None, None]
        x5 = xf.permute(3, 0, 1, 2)[None]
        acc = F.conv3d(x5, wt, padding=(1, 0, 0))[0].permute(1, 2, 3, 0).reshape(-1, N)
    if bias is not None:
        acc = acc + bias.float()
    M = acc.shape[0]
    if epilogue == EPI_GEGLU:
        v, g = _deinterleave(acc.t())
        val = (v * F.gelu(g)).t()
        val = alpha * val
        if res is not None:
            val = val + beta * res.float().reshape(M, -1)
    elif epilogue == EPI_SPADE:
        gm, bt = _deinterleave(acc.t())
        gm, bt = gm.t(), bt.t()
        Cn = gm.shape[1]
        T = shp[0]
        hv = h.float().reshape(T, -1, Cn)
        cpg = Cn // groups
        mean = gn_stats[:, :, 0].repeat_interleave(cpg, dim=1)[:, None, :]
        rstd = gn_stats[:, :, 1].repeat_interleave(cpg, dim=1)[:, None, :]
        xn = ((hv - mean) * rstd * gn_weight + gn_bias).reshape(M, Cn)
        val = xn * (1 + gm) + bt
        if res is not None:
            val = val + beta * res.float().reshape(M, -1)
    else:
        val = alpha * _act(acc, act)
        if res is not None:
            val = val + beta * res.float().reshape(M, -1)
    n_out = val.shape[1]
    dt = torch.float32 if out_f32 else torch.float16
    if stats_out is not None:
        vr = val.half().double().reshape(shp[0] if len(shp) >= 2 else 1, -1, stats_groups, n_out // stats_groups)
        stats_out += torch.stack([vr.sum(dim=(1, 3)), (vr * vr).sum(dim=(1, 3))], dim=-1)
    if out is None:
        return val.to(dt).reshape(*shp, n_out)
    out[..., out_col0:out_col0 + n_out] = val.to(out.dtype).reshape(*out.shape[:-1], n_out)
    return out


def attention(q, k, v, *, batch, heads, head_dim, nq, nkv, scale, q_col0=0, k_col0=0, v_col0=0, q_head_stride=None,
              k_head_stride=None, v_head_stride=None, kv_batched=True, out=None):
    qs = head_dim if q_head_stride is None else q_head_stride
    ks = head_dim if k_head_stride is None else k_head_stride
    vs = head_dim if v_head_stride is None else v_head_stride
    res = torch.empty(batch * nq, heads * head_dim, dtype=torch.float16) if out is None else out
    for b in range(batch):
        kb = b if kv_batched else 0
        for hh in range(heads):
            Q = q[b * nq:(b + 1) * nq, q_col0 + hh * qs: q_col0 + hh * qs + head_dim].float()
            K = k[kb * nkv:(kb + 1) * nkv, k_col0 + hh * ks: k_col0 + hh * ks + head_dim].float()
            V = v[kb * nkv:(kb + 1) * nkv, v_col0 + hh * vs: v_col0 + hh * vs + head_dim].float()
            P = torch.softmax(Q @ K.t() * scale, dim=-1)
            res[b * nq:(b + 1) * nq, hh * head_dim:(hh + 1) * head_dim] = (P @ V).half()
    return res


def _cat(x1, x2):
    return x1 if x2 is None else torch.cat([x1, x2], dim=-1)


FUSED_CONV_STATS = False


def conv_stats_slot(T, HW, device, groups=32):
    if not FUSED_CONV_STATS or HW <= 256:
        return None
    return torch.zeros(T, groups, 2, dtype=torch.float64)


def stats_pool_reset():
    pass


def stats_pool_mark_dirty():
    pass


def gn_stats(x1, x2=None, groups=32):
    x = _cat(x1, x2).double()
    T, C = x.shape[0], x.shape[-1]
    g = x.reshape(T, -1, groups, C // groups)
    return torch.stack([g.sum(dim=(1, 3)), (g * g).sum(dim=(1, 3))], dim=-1)


def _mean_rstd(sums, count, eps):
    mean = sums[..., 0] / count
    var = (sums[..., 1] / count - mean * mean).clamp_min(0)
    return mean.float(), (1.0 / torch.sqrt(var + eps)).float()


def gn_finalize(sums, HW, C, eps):
    mean, rstd = _mean_rstd(sums, HW * (C // sums.shape[1]), eps)
    return torch.stack([mean, rstd], dim=-1)


def gn_apply(x1, sums, eps, gamma, beta, silu, x2=None, groups=32):
    x = _cat(x1, x2).float()
    T, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (T * C)
    mean, rstd = _mean_rstd(sums, HW * (C // groups), eps)
    cpg = C // groups
    shp = [T] + [1] * (x.dim() - 2) + [C]
    y = (x - mean.repeat_interleave(cpg, 1).reshape(shp)) * rstd.repeat_interleave(cpg, 1).reshape(shp)
    if gamma is not None:
        y = y * gamma + beta
    if silu:
        y = F.silu(y)
    return y.half()


def group_norm(x1, gamma, beta, eps, silu, x2=None, groups=32, want_out=True, want_stats=False):
    sums = gn_stats(x1, x2, groups)
    T = x1.shape[0]
    C = x1.shape[-1] + (x2.shape[-1] if x2 is not None else 0)
    HW = x1.numel() // (T * x1.shape[-1])
    out = gn_apply(x1, sums, eps, gamma, beta, silu, x2=x2, groups=groups) if want_out else None
    stats = gn_finalize(sums, HW, C, eps) if want_stats else None
    if want_out and want_stats:
        return out, stats
    return out if want_out else stats


def layernorm(x, gamma, beta, eps=1e-5):
    return F.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps).half()


def softmax_rows(s, scale, ldp=None):
    p = torch.softmax(s.float() * scale, dim=-1).half()
    if ldp is not None and ldp != s.shape[1]:
        p = F.pad(p, (0, ldp - s.shape[1]))
    return p


def nchw_to_nhwc(x, scale=1.0):
    return (x.float() * scale).permute(0, 2, 3, 1).contiguous().half()


def nhwc_to_nchw(x, scale=1.0):
    return (x.float() * scale).permute(0, 3, 1, 2).contiguous()


def upsample2x(x):
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).contiguous()


def im2col_s2(x, pad):
    t, h, w, c = x.shape
    xp = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1) if pad == 1 else (0, 1, 0, 1))
    cols = F.unfold(xp, kernel_size=3, stride=2)  # [t, c*9, L] ordered (c, ky, kx)
    ho = (xp.shape[2] - 3) // 2 + 1
    wo = (xp.shape[3] - 3) // 2 + 1
    cols = cols.reshape(t, c, 9, ho, wo).permute(0, 3, 4, 2, 1).reshape(t, ho, wo, 9 * c)
    return cols.half().contiguous()


def conv_small_cin(x, w, bias):
    return F.conv2d(x.float(), w.float(), bias, padding=w.shape[-1] // 2).permute(0, 2, 3, 1).contiguous().half()


def conv_small_f32(x, w, bias):
    return F.conv2d(x.float(), w.float(), bias, padding=w.shape[-1] // 2)


def conv3x3_small_cout(x, w_packed, bias):
    cout = w_packed.shape[0]
    c = x.shape[-1]
    wt = w_packed.float().reshape(cout, 3, 3, c).permute(0, 3, 1, 2)
    return F.conv2d(x.float().permute(0, 3, 1, 2), wt, bias, padding=1)


def gemv(x, w, bias=None, add=None, silu_in=False, silu_out=False):
    xx = F.silu(x.float()) if silu_in else x.float()
    y = w.float() @ xx
    if bias is not None: y = y + bias
    if add is not None: y = y + add
    return F.silu(y) if silu_out else y


def timestep_embedding(t, dim, max_period=10000.0):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    a = float(t.reshape(-1)[0]) * freqs
    return torch.cat([torch.cos(a), torch.sin(a)])


def temporal_attention(qkv, heads, scale):
    t, hw, c3 = qkv.shape
    c = c3 // 3
    q, k, v = [z.float().reshape(t, hw, heads, c // heads).permute(1, 2, 0, 3) for z in qkv.split(c, dim=-1)]
    p = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1)
    o = (p @ v).permute(2, 0, 1, 3).reshape(t, hw, c)
    return o.half()


def gaussian_sample(moments, noise, scale):
    mean, logvar = torch.chunk(moments, 2, dim=1)
    std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
    return (mean + std * noise if noise is not None else mean) * scale


def axpby(x, y, a, b, relu=False):
    r = a * x.float() + b * y.float()
    return (F.relu(r) if relu else r).half()


def conv_direct(x, w, bias, stride=1, pad=0, relu=False):
    y = F.conv2d(x.float(), w.float(), bias, stride=stride, padding=pad)
    return (F.relu(y) if relu else y).permute(0, 2, 3, 1).contiguous().half()


def subsample2(x):
    return x[:, ::2, ::2].contiguous()


def instance_norm(x, relu=False, eps=1e-5):
    y = F.instance_norm(x.float().permute(0, 3, 1, 2), eps=eps)
    return (F.relu(y) if relu else y).permute(0, 2, 3, 1).contiguous().half()


def avgpool2_f32(x):
    return F.avg_pool2d(x[:, None], 2, stride=2)[:, 0]


def corr_lookup(levels, coords, out):
    from oracle.torch_ref import raft_corr_lookup
    b, _, h, w = coords.shape
    c = raft_corr_lookup([l[:, None] for l in levels], coords)          # (B, 324, h, w)
    out[..., :324] = c.permute(0, 2, 3, 1).half()
    return out


def gru_rh(zr, net):
    c = net.shape[-1]
    return (zr.float().reshape(-1, 2 * c)[:, c:].reshape(net.shape) * net.float()).half()


def gru_update(zr, q, net):
    c = net.shape[-1]
    z = zr.float().reshape(-1, 2 * c)[:, :c].reshape(net.shape)
    net.copy_(((1 - z) * net.float() + z * q.float()).half())
    return net


def set_channels(src, dst, col0):
    dst[..., col0:col0 + src.shape[1]] = src.permute(0, 2, 3, 1).half()
    return dst


def convex_upsample8(mask, flow):
    from oracle.torch_ref import raft_upsample_flow
    return raft_upsample_flow(flow, mask.float().permute(0, 3, 1, 2))


# ---- flow ops: emulated with the same torch ops the reference uses -------------------------------------------------
def _grid(h, w, flow_hw2):
    gy, gx = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
    vg = torch.stack((gx, gy), 2)[None] + flow_hw2
    return torch.stack((2.0 * vg[..., 0] / max(w - 1, 1) - 1.0, 2.0 * vg[..., 1] / max(h - 1, 1) - 1.0), dim=3)


def flow_warp_f32(x, flow, flow_layout=0, nearest=False, border=False, align_corners=True):
    fl = flow if flow_layout == 0 else flow.permute(0, 2, 3, 1)
    return F.grid_sample(x, _grid(x.shape[2], x.shape[3], fl), mode="nearest" if nearest else "bilinear",
                         padding_mode="border" if border else "zeros", align_corners=align_corners)


def fb_consistency_f32(fwd_flow, bwd_flow, alpha=0.01, beta=0.5):
    from oracle.torch_ref import forward_backward_consistency_check
    return forward_backward_consistency_check(fwd_flow, bwd_flow, alpha, beta)


def motion_guidance_f32(latents, flow_fwd_prop, flow_bwd_prop, fwd_occ, bwd_occ, step, want_loss=False):
    from oracle.torch_ref import temporal_condition_v4
    t = latents.shape[0]
    with torch.enable_grad():
        lat = latents.detach().clone().requires_grad_(True)
        loss = temporal_condition_v4((flow_fwd_prop[None], flow_bwd_prop[None]), lat,
                                     (fwd_occ[None, :, None], bwd_occ[None, :, None]), t)
        g = torch.autograd.grad(loss, lat)[0] if torch.is_tensor(loss) else torch.zeros_like(latents)
    out = (latents - step * g).detach()
    return (out, loss.detach().reshape(1), g) if want_loss else out


def resize_flow_f32(flow, oh, ow):
    from oracle.torch_ref import resize_flow
    return resize_flow(flow, oh, ow)


def canvas_posterior_f32(x, eps_tiles, tile_w, noise, offsets, tile_size, c_recip, c_recipm1, c1, c2, sigma,
                         want_eps=False):
    acc = torch.zeros_like(x)
    cnt = torch.zeros_like(x)
    for e, (ox, oy) in zip(eps_tiles, offsets):
        acc[:, :, oy:oy + tile_size, ox:ox + tile_size] += e * tile_w
        cnt[:, :, oy:oy + tile_size, ox:ox + tile_size] += tile_w
    eps = acc / cnt
    x0 = c_recip * x - c_recipm1 * eps
    mean = c1 * x0 + c2 * x
    out = mean + sigma * noise if noise is not None else mean
    return (out, eps) if want_eps else out


def frames_u8_to_f32_bicubic(frames_u8, oh, ow, pad_h=0, pad_w=0, clamp=False):
    x = (frames_u8.permute(0, 3, 1, 2).float() / 255.0 - 0.5) / 0.5
    x = F.interpolate(x, size=(oh, ow), mode="bicubic")
    if clamp:
        x = x.clamp(-1.0, 1.0)
    if pad_h or pad_w:
        x = F.pad(x, (0, pad_w, 0, pad_h), mode="reflect")
    return x


def frames_f32_to_u8_hwc(frames, crop_h=None, crop_w=None):
    n, c, h, w = frames.shape
    x = (frames.permute(0, 2, 3, 1) * 255.0)[:, :crop_h or h, :crop_w or w]
    return torch.from_numpy(x.cpu().numpy().astype("uint8")).to(frames.device)
