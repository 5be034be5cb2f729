"""TEST INFRASTRUCTURE — CPU/torch restatement of the MGLD-VSR hot path (the parity oracle).

This file is NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs import it.  It restates the reference's algorithm for the hot path in
plain functional PyTorch (fp32, NCHW, standard ``torch.nn.functional`` ops), driven directly by a reference
``state_dict`` — every function cites the reference file:line it follows (paths relative to the reference root).

Pinning: the reference has no golden vectors or tests on this path.  The restatement is pinned against
the reference's *own modules* imported from /root/reference through ``oracle/ref_shim.py`` (tests/test_oracle.py and,
for the sampler and the script's segment loop, tests/test_reference_pipeline.py; CPU, marker ``reference``, run in the
build container) and against the fixtures in ``tests/golden/`` that were generated from those modules by
``tools/make_golden.py``.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------------
# small helpers
# ---------------------------------------------------------------------------------------------------------------
def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _gn(sd, p, x, eps):
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps).type(x.dtype)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def attention(q, k, v, scale=None):
    """xformers.ops.memory_efficient_attention semantics: exact softmax(q k^T * scale) v; default scale d^-0.5."""
    scale = q.shape[-1] ** -0.5 if scale is None else scale
    w = torch.softmax(torch.einsum("bid,bjd->bij", q.float(), k.float()) * scale, dim=-1)
    return torch.einsum("bij,bjd->bid", w, v.float()).type(q.dtype)


def timestep_embedding(timesteps, dim, max_period=10000):
    """ldm/modules/diffusionmodules/util.py:151-171"""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _split_heads(t, heads):
    b, n, c = t.shape
    return t.reshape(b, n, heads, c // heads).permute(0, 2, 1, 3).reshape(b * heads, n, c // heads)


def _merge_heads(t, heads):
    bh, n, d = t.shape
    return t.reshape(bh // heads, heads, n, d).permute(0, 2, 1, 3).reshape(bh // heads, n, heads * d)


# ---------------------------------------------------------------------------------------------------------------
# UNet pieces
# ---------------------------------------------------------------------------------------------------------------
def spatial_temporal_conv(sd, p, x, num_frames):
    """SpatialTemporalConv.forward, util.py:301-310 (with oracle patch D1: t = num_frames)."""
    bt, c, h, w = x.shape
    b = bt // num_frames
    x5 = x.reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
    res = F.conv3d(x5, sd[p + ".temporal_conv.weight"], sd[p + ".temporal_conv.bias"], padding=(1, 0, 0))
    res = res.permute(0, 2, 1, 3, 4).reshape(bt, c, h, w)
    a = sd[p + ".temporal_alpha"]
    return a * res + (1 - a) * x


def self_attention(sd, p, x, heads, context=None):
    """MemoryEfficientSelfAttention / MemoryEfficientCrossAttention.forward, attention.py:283-309, 333-381."""
    q = _lin(sd, p + ".to_q", x)
    ctx = x if context is None else context
    if ctx.shape[0] != x.shape[0]:
        ctx = torch.repeat_interleave(ctx, x.shape[0] // ctx.shape[0], dim=0)  # attention.py:336-337
    k = _lin(sd, p + ".to_k", ctx)
    v = _lin(sd, p + ".to_v", ctx)
    out = attention(_split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads))
    return _lin(sd, p + ".to_out.0", _merge_heads(out, heads))


def temporal_attention(sd, p, x, heads, num_frames):
    """TemporalAttention.forward, attention.py:135-143 (oracle patch D1)."""
    bt, c, h, w = x.shape
    b = bt // num_frames
    seq = x.reshape(b, num_frames, c, h, w).permute(0, 3, 4, 1, 2).reshape(b * h * w, num_frames, c)
    res = self_attention(sd, p + ".temporal_attn", _ln(sd, p + ".norm", seq), heads)
    res = res.reshape(b, h, w, num_frames, c).permute(0, 3, 4, 1, 2).reshape(bt, c, h, w)
    a = sd[p + ".temporal_alpha"]
    return a * res + (1 - a) * x


def spatial_transformer(sd, p, x, context, heads):
    """SpatialTransformerV2.forward (use_linear=True, depth 1), attention.py:527-546 + BasicTransformerBlockV2 :431-435."""
    b, c, h, w = x.shape
    x_in = x
    t = F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    t = t.reshape(b, c, h * w).permute(0, 2, 1)
    t = _lin(sd, p + ".proj_in", t)
    bp = p + ".transformer_blocks.0"
    t = self_attention(sd, bp + ".attn1", _ln(sd, bp + ".norm1", t), heads) + t
    t = self_attention(sd, bp + ".attn2", _ln(sd, bp + ".norm2", t), heads, context=context) + t
    hdn = _lin(sd, bp + ".ff.net.0.proj", _ln(sd, bp + ".norm3", t))
    val, gate = hdn.chunk(2, dim=-1)
    t = _lin(sd, bp + ".ff.net.2", val * F.gelu(gate)) + t
    t = _lin(sd, p + ".proj_out", t)
    return t.permute(0, 2, 1).reshape(b, c, h, w) + x_in


def spade(sd, p, x, segmap_dic):
    """SPADE.forward, spade.py:90-111"""
    seg = segmap_dic[str(x.shape[-1])]
    normalized = _gn(sd, p + ".param_free_norm", x, 1e-5)
    actv = F.relu(_conv(sd, p + ".mlp_shared.0", seg))
    return normalized * (1 + _conv(sd, p + ".mlp_gamma", actv)) + _conv(sd, p + ".mlp_beta", actv)


def res_block(sd, p, x, emb, s_cond=None):
    """ResBlock._forward openaimodel.py:335-359 / ResBlockDual._forward :459-482 (no up/down, no scale-shift)."""
    h = _conv(sd, p + ".in_layers.2", F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5)))
    emb_out = _lin(sd, p + ".emb_layers.1", F.silu(emb)).type(h.dtype)
    h = h + emb_out[:, :, None, None]
    h = _conv(sd, p + ".out_layers.3", F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5)))
    if s_cond is not None:
        h = spade(sd, p + ".spade", h, s_cond)
    if p + ".skip_connection.weight" in sd:
        x = _conv(sd, p + ".skip_connection", x, padding=0)
    return x + h


def unet_layout(cfg):
    """Block structure of InflatedUNetModelDualcondV2.__init__ (openaimodel.py:2036-2257) as a list of layer kinds."""
    mc, mult, nrb = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    attn_res, nhc = cfg["attention_resolutions"], cfg["num_head_channels"]
    inp, chans, ch, ds = [["conv_in"]], [mc], mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", ch, m * mc)]
            ch = m * mc
            if ds in attn_res:
                layers.append(("st", ch, ch // nhc))
            inp.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    mid = [("res", ch, ch), ("stconv", ch), ("st", ch, ch // nhc), ("tattn", ch, ch // nhc), ("res", ch, ch),
           ("stconv", ch)]
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * m)]
            ch = mc * m
            if ds in attn_res:
                layers.append(("st", ch, ch // nhc))
            if level and i == nrb:
                layers.append(("up", ch))
                ds //= 2
            out.append(layers)
    return inp, mid, out


def _run_layers(sd, p, layers, h, emb, context, struct_cond, num_frames):
    for j, l in enumerate(layers):
        q = f"{p}.{j}"
        if l[0] == "res":
            h = res_block(sd, q, h, emb, struct_cond)
        elif l[0] == "st":
            h = spatial_transformer(sd, q, h, context, l[2])
        elif l[0] == "stconv":
            h = spatial_temporal_conv(sd, q, h, num_frames)
        elif l[0] == "tattn":
            h = temporal_attention(sd, q, h, l[2], num_frames)
        elif l[0] == "down":
            h = _conv(sd, q + ".op", h, stride=2, padding=1)                       # openaimodel.py:204-231
        elif l[0] == "up":
            h = _conv(sd, q + ".conv", F.interpolate(h, scale_factor=2, mode="nearest"))  # openaimodel.py:178-188
        elif l == "conv_in":
            h = _conv(sd, q, h)
    return h


def unet_forward(sd, cfg, x, timesteps, context, struct_cond, prefix="model.diffusion_model."):
    """InflatedUNetModelDualcondV2.forward, openaimodel.py:2281-2313.  sd keys carry `prefix`."""
    sd = _strip(sd, prefix)
    nf = cfg["num_frames"]
    emb = _lin(sd, "time_embed.2", F.silu(_lin(sd, "time_embed.0", timestep_embedding(timesteps, cfg["model_channels"]))))
    inp, mid, out = unet_layout(cfg)
    hs, h = [], x
    for i, layers in enumerate(inp):
        h = _run_layers(sd, f"input_blocks.{i}", layers, h, emb, context, struct_cond, nf)
        hs.append(h)
    h = _run_layers(sd, "middle_block", mid, h, emb, context, struct_cond, nf)
    for i, layers in enumerate(out):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(sd, f"output_blocks.{i}", layers, h, emb, context, struct_cond, nf)
    return _conv(sd, "out.2", F.silu(_gn(sd, "out.0", h, 1e-5)))


# ---------------------------------------------------------------------------------------------------------------
# struct-cond encoder
# ---------------------------------------------------------------------------------------------------------------
def attention_block(sd, p, x, heads):
    """AttentionBlock._forward + QKVAttentionLegacy (xformers branch), openaimodel.py:525-531, 565-590."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(sd, p + ".norm", xf, 1e-5), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    ch = c // heads
    q, k, v = qkv.reshape(b * heads, ch * 3, -1).split(ch, dim=1)
    a = attention(q.permute(0, 2, 1), k.permute(0, 2, 1), v.permute(0, 2, 1)).permute(0, 2, 1).reshape(b, c, -1)
    hproj = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + hproj).reshape(b, c, hh, ww)


def struct_layout(cfg):
    """InflatedEncoderUNetModelWT.__init__ block structure, openaimodel.py:2372-2484."""
    mc, mult, nrb = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    attn_res = cfg["attention_resolutions"]
    blocks, chans, ch, ds = [["conv_in"]], [], mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", ch, m * mc)]
            ch = m * mc
            if ds in attn_res:
                layers.append(("attn", ch))
            blocks.append(layers)
        if level != len(mult) - 1:
            blocks.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    chans.append(ch)
    return blocks, ch, chans


def struct_encoder_forward(sd, cfg, x, timesteps, prefix="structcond_stage_model."):
    """InflatedEncoderUNetModelWT.forward, openaimodel.py:2500-2525."""
    sd = _strip(sd, prefix)
    heads = cfg["num_heads"]
    emb = _lin(sd, "time_embed.2", F.silu(_lin(sd, "time_embed.0", timestep_embedding(timesteps, cfg["model_channels"]))))
    blocks, ch, chans = struct_layout(cfg)
    results, h = [], x
    for i, layers in enumerate(blocks):
        last = h
        for j, l in enumerate(layers):
            q = f"input_blocks.{i}.{j}"
            if l == "conv_in":
                h = _conv(sd, q, h)
            elif l[0] == "res":
                h = res_block(sd, q, h, emb)
            elif l[0] == "attn":
                h = attention_block(sd, q, h, heads)
            elif l[0] == "down":
                h = _conv(sd, q + ".op", h, stride=2, padding=1)
        if h.shape[-1] != last.shape[-1]:
            results.append(last)
    h = res_block(sd, "middle_block.0", h, emb)
    h = attention_block(sd, "middle_block.1", h, heads)
    h = res_block(sd, "middle_block.2", h, emb)
    results.append(h)
    return {str(r.shape[-1]): res_block(sd, f"fea_tran.{i}", r, emb) for i, r in enumerate(results)}


# ---------------------------------------------------------------------------------------------------------------
# VAE
# ---------------------------------------------------------------------------------------------------------------
def _swish(x):
    return x * torch.sigmoid(x)


def _gn6(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def vae_resnet_block(sd, p, x):
    """ResnetBlock.forward (temb None), model.py:163-183"""
    h = _conv(sd, p + ".conv1", _swish(_gn6(sd, p + ".norm1", x)))
    h = _conv(sd, p + ".conv2", _swish(_gn6(sd, p + ".norm2", h)))
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(sd, p + ".nin_shortcut", x, padding=0)
    return x + h


def vae_attn_block(sd, p, x):
    """MemoryEfficientAttnBlock.forward, model.py:274-305: single head, head-dim C, scale C^-0.5"""
    b, c, h, w = x.shape
    hn = _gn6(sd, p + ".norm", x)
    q, k, v = [_conv(sd, f"{p}.{n}", hn, padding=0).reshape(b, c, h * w).permute(0, 2, 1) for n in "qkv"]
    out = attention(q, k, v, scale=int(c) ** -0.5).permute(0, 2, 1).reshape(b, c, h, w)
    return x + _conv(sd, p + ".proj_out", out, padding=0)


def vae_encoder_forward(sd, ddcfg, x, prefix="encoder.", return_fea=False):
    """Encoder.forward, model.py:539-572 (feature taps after levels 1 and 2, :552-554)."""
    sd = _strip(sd, prefix)
    nres, nrb = len(ddcfg["ch_mult"]), ddcfg["num_res_blocks"]
    h = _conv(sd, "conv_in", x)
    fea = []
    for lvl in range(nres):
        for b in range(nrb):
            h = vae_resnet_block(sd, f"down.{lvl}.block.{b}", h)
        if return_fea and lvl in (1, 2):
            fea.append(h)
        if lvl != nres - 1:
            h = _conv(sd, f"down.{lvl}.downsample.conv", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)  # model.py:114-117
    h = vae_resnet_block(sd, "mid.block_1", h)
    h = vae_attn_block(sd, "mid.attn_1", h)
    h = vae_resnet_block(sd, "mid.block_2", h)
    h = _conv(sd, "conv_out", _swish(_gn6(sd, "norm_out", h)))
    return (h, fea) if return_fea else h


def autoencoder_kl_encode(sd, ddcfg, x, prefix="first_stage_model."):
    """AutoencoderKL.encode, autoencoder.py:347-353 -> moments (mean | logvar)."""
    h = vae_encoder_forward(sd, ddcfg, x, prefix + "encoder.")
    return F.conv2d(h, sd[prefix + "quant_conv.weight"], sd[prefix + "quant_conv.bias"])


def gaussian_sample(moments, noise):
    """DiagonalGaussianDistribution.sample, distributions.py:24-37 (noise supplied by the caller)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise


def _fuse_resblock(sd, p, x):
    """model.py:1312-1335 ResBlock (fusion layers)."""
    h = _conv(sd, p + ".conv1", _swish(_gn6(sd, p + ".norm1", x)))
    h = _conv(sd, p + ".conv2", _swish(_gn6(sd, p + ".norm2", h)))
    if p + ".conv_out.weight" in sd:
        x = _conv(sd, p + ".conv_out", x, padding=0)
    return h + x


def _rdb(sd, p, x):
    """basicsr/archs/rrdbnet_arch.py:32-39 ResidualDenseBlock.forward"""
    lr = lambda t: F.leaky_relu(t, 0.2)
    x1 = lr(_conv(sd, p + ".conv1", x))
    x2 = lr(_conv(sd, p + ".conv2", torch.cat((x, x1), 1)))
    x3 = lr(_conv(sd, p + ".conv3", torch.cat((x, x1, x2), 1)))
    x4 = lr(_conv(sd, p + ".conv4", torch.cat((x, x1, x2, x3), 1)))
    x5 = _conv(sd, p + ".conv5", torch.cat((x, x1, x2, x3, x4), 1))
    return x5 * 0.2 + x


def fuse_block(sd, p, enc_feat, dec_feat, w, num_block=2):
    """Fuse_sft_block_ResidualDenseBlock.forward, model.py:1361-1367"""
    e = _fuse_resblock(sd, p + ".encode_enc_1", torch.cat([enc_feat, dec_feat], dim=1))
    for i in range(num_block):
        e = _rdb(sd, f"{p}.encode_enc_2.{i}", e)
    e = _fuse_resblock(sd, p + ".encode_enc_3", e)
    return dec_feat + w * e


def video_decoder_forward(sd, ddcfg, z, enc_fea, fusion_w=1.0, prefix="decoder."):
    """VideoDecoder_Mix.forward, model.py:1017-1056."""
    sd = _strip(sd, prefix)
    nres, nrb, nf = len(ddcfg["ch_mult"]), ddcfg["num_res_blocks"], ddcfg["num_frames"]
    h = _conv(sd, "conv_in", z)
    h = vae_resnet_block(sd, "mid.block_1", h)
    h = spatial_temporal_conv(sd, "temporal_mixing", h, nf)
    h = vae_attn_block(sd, "mid.attn_1", h)
    h = vae_resnet_block(sd, "mid.block_2", h)
    for lvl in reversed(range(nres)):
        for b in range(nrb + 1):
            h = vae_resnet_block(sd, f"up.{lvl}.block.{b}", h)
            h = spatial_temporal_conv(sd, f"up.{lvl}.temporal_mixing.{b}", h, nf)
        if lvl != nres - 1 and lvl != 0:
            h = fuse_block(sd, f"fusion_layer_{lvl}", enc_fea[lvl - 1], h, fusion_w)
        if lvl != 0:
            h = _conv(sd, f"up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    return _conv(sd, "conv_out", _swish(_gn6(sd, "norm_out", h)))


def decoder_forward(sd, ddcfg, z, prefix="decoder."):
    """Decoder.forward (plain image decoder of AutoencoderKL), model.py:648-684."""
    sd = _strip(sd, prefix)
    nres, nrb = len(ddcfg["ch_mult"]), ddcfg["num_res_blocks"]
    h = _conv(sd, "conv_in", z)
    h = vae_resnet_block(sd, "mid.block_1", h)
    h = vae_attn_block(sd, "mid.attn_1", h)
    h = vae_resnet_block(sd, "mid.block_2", h)
    for lvl in reversed(range(nres)):
        for b in range(nrb + 1):
            h = vae_resnet_block(sd, f"up.{lvl}.block.{b}", h)
        if lvl != 0:
            h = _conv(sd, f"up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    return _conv(sd, "conv_out", _swish(_gn6(sd, "norm_out", h)))


def autoencoder_kl_decode(sd, ddcfg, z, prefix="first_stage_model."):
    """AutoencoderKL.decode, autoencoder.py:361-364"""
    z = F.conv2d(z, sd[prefix + "post_quant_conv.weight"], sd[prefix + "post_quant_conv.bias"])
    return decoder_forward(sd, ddcfg, z, prefix + "decoder.")


def video_vae_encode(sd, ddcfg, x):
    """VideoAutoencoderKLResi.encode, autoencoder.py:1674-1679 -> (moments, enc_fea)"""
    h, fea = vae_encoder_forward(sd, ddcfg, x, "encoder.", return_fea=True)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"]), fea


def video_vae_decode(sd, ddcfg, z, enc_fea, fusion_w=1.0):
    """VideoAutoencoderKLResi.decode, autoencoder.py:1687-1690"""
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    return video_decoder_forward(sd, ddcfg, z, enc_fea, fusion_w)


class _Stripped(dict):
    pass


def _strip(sd, prefix):
    if not prefix or isinstance(sd, _Stripped) and getattr(sd, "prefix", None) == prefix:
        return sd
    out = _Stripped((k[len(prefix):], v) for k, v in sd.items() if k.startswith(prefix))
    out.prefix = prefix
    return out


# ---------------------------------------------------------------------------------------------------------------
# schedule (float64 numpy -> float32 tables, exactly as the reference registers them)
# ---------------------------------------------------------------------------------------------------------------
def make_beta_schedule_linear(n_timestep, linear_start, linear_end):
    """util.py:21-27 ('linear' = linear in sqrt(beta))"""
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


def schedule_tables(betas):
    """DDPM.register_schedule, ddpm.py:237-277 (v_posterior = 0, eps-parameterisation)."""
    betas = np.asarray(betas)   # NOT cast: the script passes float32 betas and numpy then computes the tables in float32
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "betas": f32(betas), "alphas_cumprod": f32(ac), "alphas_cumprod_prev": f32(ac_prev),
        "sqrt_alphas_cumprod": f32(np.sqrt(ac)), "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - ac)),
        "sqrt_recip_alphas_cumprod": f32(np.sqrt(1.0 / ac)), "sqrt_recipm1_alphas_cumprod": f32(np.sqrt(1.0 / ac - 1)),
        "posterior_variance": f32(post_var),
        "posterior_log_variance_clipped": f32(np.log(np.maximum(post_var, 1e-20))),
        "posterior_mean_coef1": f32(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        "posterior_mean_coef2": f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
    }


def space_timesteps(num_timesteps, section_counts):
    """scripts/vsr_val_ddpm_text_T_vqganfin_oldcanvas_tile.py:33-88 (integer list form)."""
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start_idx, all_steps = 0, []
    for i, section_count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < section_count:
            raise ValueError(f"cannot divide section of {size} steps into {section_count}")
        frac_stride = 1 if section_count <= 1 else (size - 1) / (section_count - 1)
        cur_idx, taken = 0.0, []
        for _ in range(section_count):
            taken.append(start_idx + round(cur_idx))
            cur_idx += frac_stride
        all_steps += taken
        start_idx += size
    return set(all_steps)


def respaced_schedule(linear_start=0.00085, linear_end=0.0120, timesteps=1000, ddpm_steps=50):
    """script :308-328: the 1000-step tables kept for q_sample_respace + the respaced S-step schedule."""
    base = schedule_tables(make_beta_schedule_linear(timesteps, linear_start, linear_end))
    use = space_timesteps(timesteps, [ddpm_steps])
    last, new_betas = 1.0, []
    # the script iterates the float32 `model.alphas_cumprod` buffer (script :318-323)
    for i, ac in enumerate(base["alphas_cumprod"]):
        if i in use:
            new_betas.append(1 - ac / last)
            last = ac
    new_betas = np.array([b.data.cpu().numpy() for b in new_betas])   # float32, exactly as script :324
    return base, schedule_tables(new_betas), sorted(use)


# ---------------------------------------------------------------------------------------------------------------
# flow ops + guidance
# ---------------------------------------------------------------------------------------------------------------
def flow_warp(x, flow, interp_mode="bilinear", padding_mode="zeros", align_corners=True):
    """basicsr/archs/arch_util.py:156-184; flow (n,h,w,2)"""
    _, _, h, w = x.shape
    gy, gx = torch.meshgrid(torch.arange(0, h).type_as(x), torch.arange(0, w).type_as(x), indexing="ij")
    vgrid = torch.stack((gx, gy), 2).float() + flow
    vx = 2.0 * vgrid[:, :, :, 0] / max(w - 1, 1) - 1.0
    vy = 2.0 * vgrid[:, :, :, 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(x, torch.stack((vx, vy), dim=3), mode=interp_mode, padding_mode=padding_mode,
                         align_corners=align_corners)


def resize_flow(flow, out_h, out_w):
    """arch_util.py:235-270 with size_type='shape'"""
    _, _, fh, fw = flow.shape
    inp = flow.clone()
    inp[:, 0] *= out_w / fw
    inp[:, 1] *= out_h / fh
    return F.interpolate(inp, size=(out_h, out_w), mode="bilinear", align_corners=False)


def forward_backward_consistency_check(fwd_flow, bwd_flow, alpha=0.01, beta=0.5):
    """scripts/util_flow.py:114-136"""
    def warp(feat, flow):  # util_flow.py:97-111, 64-94
        b, c, h, w = feat.shape
        y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        grid = torch.stack([x, y], 0).float()[None].to(flow.device) + flow
        xg = 2 * grid[:, 0] / (w - 1) - 1
        yg = 2 * grid[:, 1] / (h - 1) - 1
        return F.grid_sample(feat, torch.stack([xg, yg], -1), mode="bilinear", padding_mode="zeros", align_corners=True)
    mag = torch.norm(fwd_flow, dim=1) + torch.norm(bwd_flow, dim=1)
    diff_fwd = torch.norm(fwd_flow + warp(bwd_flow, fwd_flow), dim=1)
    diff_bwd = torch.norm(bwd_flow + warp(fwd_flow, bwd_flow), dim=1)
    thr = alpha * mag + beta
    return (diff_fwd > thr).float(), (diff_bwd > thr).float()


def temporal_condition_v4(flows, latents, masks, num_frames):
    """compute_temporal_condition_v4, ddpm.py:3538-3574.  flows: 2 x (b,t-1,2,h,w); masks: 2 x (b,t-1,1,h,w)."""
    flow_fwd_prop, flow_bwd_prop = flows
    fwd_occs, bwd_occs = masks
    t = num_frames
    lat = latents.reshape(-1, t, *latents.shape[1:])
    loss_b, warp, prev = 0, torch.zeros_like(lat[:, -1]), None
    for i in range(t - 1, -1, -1):
        cur = lat[:, i]
        if i < t - 1:
            warp = flow_warp(cur, flow_bwd_prop[:, i].permute(0, 2, 3, 1))
            loss_b = loss_b + F.l1_loss((1 - fwd_occs[:, i]) * prev, (1 - fwd_occs[:, i]) * cur)
        prev = warp
    loss_f, warp = 0, torch.zeros_like(lat[:, 0])
    for i in range(t):
        cur = lat[:, i]
        if i > 0:
            warp = flow_warp(cur, flow_fwd_prop[:, i - 1].permute(0, 2, 3, 1))
            loss_f = loss_f + F.l1_loss((1 - bwd_occs[:, i - 1]) * prev, (1 - bwd_occs[:, i - 1]) * cur)
        prev = warp
    return loss_b + loss_f


def guidance_update(latents, flows, masks, num_frames, guidance_scale, model_log_variance):
    """ddpm.py:4429-4435"""
    with torch.enable_grad():
        lat = latents.detach().clone().requires_grad_(True)
        loss = temporal_condition_v4(flows, lat, masks, num_frames)
        if not torch.is_tensor(loss):
            return latents
        g = torch.autograd.grad(loss, lat)[0]
    return (lat - guidance_scale * model_log_variance * g).detach()


# ---------------------------------------------------------------------------------------------------------------
# sampler
# ---------------------------------------------------------------------------------------------------------------
def gaussian_weights(tile_width, tile_height, nbatches=1):
    """LatentDiffusionVSRTextWT._gaussian_weights, ddpm.py:4601-4616 (float64, asymmetric midpoints — quirk D9)."""
    from numpy import exp, pi, sqrt
    var = 0.01
    midpoint = (tile_width - 1) / 2
    x_probs = [exp(-(x - midpoint) * (x - midpoint) / (tile_width * tile_width) / (2 * var)) / sqrt(2 * pi * var)
               for x in range(tile_width)]
    midpoint = tile_height / 2
    y_probs = [exp(-(y - midpoint) * (y - midpoint) / (tile_height * tile_height) / (2 * var)) / sqrt(2 * pi * var)
               for y in range(tile_height)]
    weights = np.outer(y_probs, x_probs)
    return torch.tile(torch.tensor(weights), (nbatches, 4, 1, 1))


def canvas_tiles(h, w, tile_size, tile_overlap):
    """tile grid of p_mean_variance_canvas, ddpm.py:4203-4231: returns [(ofs_x, ofs_y)] in (row, col) order."""
    rows, cur = 0, 0
    while cur < w:
        cur = max(rows * tile_size - tile_overlap * rows, 0) + tile_size
        rows += 1
    cols, cur = 0, 0
    while cur < h:
        cur = max(cols * tile_size - tile_overlap * cols, 0) + tile_size
        cols += 1
    out = []
    for row in range(rows):
        for col in range(cols):
            ofs_x = max(row * tile_size - tile_overlap * row, 0)
            ofs_y = max(col * tile_size - tile_overlap * col, 0)
            if row == rows - 1:
                ofs_x = w - tile_size
            if col == cols - 1:
                ofs_y = h - tile_size
            out.append((ofs_x, ofs_y))
    return out


class RefModel:
    """Functional stand-in for LatentDiffusionVSRTextWT (only what the sampling path touches)."""

    def __init__(self, sd, unet_cfg, struct_cfg, sched, ori_timesteps, num_frames):
        self.sd, self.unet_cfg, self.struct_cfg = sd, unet_cfg, struct_cfg
        self.sched, self.ori_timesteps, self.num_frames = sched, ori_timesteps, num_frames
        self.unet_sd = _strip(sd, "model.diffusion_model.")
        self.struct_sd = _strip(sd, "structcond_stage_model.")

    def eps(self, x, t_in, context, struct_cond_tile):
        sc = struct_encoder_forward(self.struct_sd, self.struct_cfg, struct_cond_tile, t_in, prefix="")
        return unet_forward(self.unet_sd, self.unet_cfg, x, t_in, context, sc, prefix="")

    def posterior(self, x, eps, i):
        """predict_start_from_noise ddpm.py:340-344 + q_posterior :346-353"""
        s = {k: v.to(x.device) for k, v in self.sched.items()}
        x0 = s["sqrt_recip_alphas_cumprod"][i] * x - s["sqrt_recipm1_alphas_cumprod"][i] * eps
        mean = s["posterior_mean_coef1"][i] * x0 + s["posterior_mean_coef2"][i] * x
        return mean, s["posterior_log_variance_clipped"][i]

    def p_sample_canvas(self, x, context, struct_cond, i, noise, flows, masks, guidance_scale, tile_size, tile_overlap,
                        tile_weights):
        """p_sample_canvas ddpm.py:4383-4440 + p_mean_variance_canvas :4191-4322 (batch_size_sample = 1)."""
        t_in = torch.full((1,), self.ori_timesteps[i], dtype=torch.long, device=x.device)
        _, _, h, w = x.shape
        noise_pred = torch.zeros(x.shape, device=x.device)
        contributors = torch.zeros(x.shape, device=x.device)
        for (ox, oy) in canvas_tiles(h, w, tile_size, tile_overlap):
            e = self.eps(x[:, :, oy:oy + tile_size, ox:ox + tile_size], t_in, context,
                         struct_cond[:, :, oy:oy + tile_size, ox:ox + tile_size])
            noise_pred[:, :, oy:oy + tile_size, ox:ox + tile_size] += e * tile_weights
            contributors[:, :, oy:oy + tile_size, ox:ox + tile_size] += tile_weights
        noise_pred /= contributors
        mean, logvar = self.posterior(x, noise_pred, i)
        nonzero = 0.0 if i == 0 else 1.0
        lat = mean + nonzero * (0.5 * logvar).exp() * noise
        if flows is not None:
            lat = guidance_update(lat, flows, masks, self.num_frames, guidance_scale, logvar)
        return lat, noise_pred

    def sample_canvas(self, context, struct_cond, x_T, noises, flows=None, masks=None, guidance_scale=-10.0,
                      tile_size=64, tile_overlap=32, return_eps=False):
        """sample_canvas / p_sample_loop_canvas, ddpm.py:4722-4760, 4619-4694.  noises[i] = the randn drawn at step i."""
        S = len(self.ori_timesteps)
        tw = gaussian_weights(tile_size, tile_size, 1).to(x_T.device)
        img, eps_trace = x_T, []
        for i in reversed(range(S)):
            img, e = self.p_sample_canvas(img, context, struct_cond, i, noises[i], flows, masks, guidance_scale,
                                          tile_size, tile_overlap, tw)
            if return_eps:
                eps_trace.append(e)
        return (img, eps_trace) if return_eps else img


# ---------------------------------------------------------------------------------------------------------------
# RAFT_SR ("normal" model) — basicsr/archs/raft_arch.py
# ---------------------------------------------------------------------------------------------------------------
def _raft_norm(sd, p, x, kind):
    if kind == "instance":
        return F.instance_norm(x)                                       # nn.InstanceNorm2d defaults: eps 1e-5, no affine
    if kind == "batch":
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            False, 0.0, 1e-5)
    return x


def _raft_resblock(sd, p, x, kind, stride):
    """ResidualBlock.forward, raft_arch.py:127-136"""
    y = F.relu(_raft_norm(sd, p + ".norm1", _conv(sd, p + ".conv1", x, stride=stride), kind))
    y = F.relu(_raft_norm(sd, p + ".norm2", _conv(sd, p + ".conv2", y), kind))
    if stride != 1:
        x = _raft_norm(sd, p + ".norm3", _conv(sd, p + ".downsample.0", x, stride=stride, padding=0), kind)
    return F.relu(x + y)


def raft_encoder(sd, p, x, kind):
    """BasicEncoder.forward, raft_arch.py:247-268"""
    x = F.relu(_raft_norm(sd, p + ".norm1", _conv(sd, p + ".conv1", x, stride=2, padding=3), kind))
    for li, stride in ((1, 1), (2, 2), (3, 2)):
        x = _raft_resblock(sd, f"{p}.layer{li}.0", x, kind, stride)
        x = _raft_resblock(sd, f"{p}.layer{li}.1", x, kind, 1)
    return _conv(sd, p + ".conv2", x, padding=0)


def _raft_bilinear_sampler(img, coords):
    """raft_arch.py:517-532"""
    H, W = img.shape[-2:]
    xg, yg = coords.split([1, 1], dim=-1)
    grid = torch.cat([2 * xg / (W - 1) - 1, 2 * yg / (H - 1) - 1], dim=-1)
    return F.grid_sample(img, grid, align_corners=True)


def raft_corr_pyramid(fmap1, fmap2, num_levels=4):
    """CorrBlock.__init__ + CorrBlock.corr, raft_arch.py:37-55, 83-92"""
    b, dim, ht, wd = fmap1.shape
    corr = torch.matmul(fmap1.view(b, dim, ht * wd).transpose(1, 2), fmap2.view(b, dim, ht * wd))
    corr = corr.view(b, ht, wd, 1, ht, wd) / torch.sqrt(torch.tensor(dim).float())
    corr = corr.reshape(b * ht * wd, 1, ht, wd)
    pyr = [corr]
    for _ in range(num_levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    return pyr


def raft_corr_lookup(pyr, coords, radius=4):
    """CorrBlock.__call__, raft_arch.py:57-81"""
    r = radius
    coords = coords.permute(0, 2, 3, 1)
    b, h1, w1, _ = coords.shape
    out = []
    for i, corr in enumerate(pyr):
        dx = torch.linspace(-r, r, 2 * r + 1)
        dy = torch.linspace(-r, r, 2 * r + 1)
        delta = torch.stack(torch.meshgrid(dy, dx, indexing="ij"), axis=-1).to(coords.device)
        centroid = coords.reshape(b * h1 * w1, 1, 1, 2) / 2 ** i
        c = _raft_bilinear_sampler(corr, centroid + delta.view(1, 2 * r + 1, 2 * r + 1, 2))
        out.append(c.view(b, h1, w1, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def _raft_update(sd, net, inp, corr, flow, p="update_block"):
    """BasicUpdateBlock.forward raft_arch.py:476-486 with BasicMotionEncoder :436-445, SepConvGRU :398-412, FlowHead"""
    e = p + ".encoder"
    cor = F.relu(_conv(sd, e + ".convc1", corr, padding=0))
    cor = F.relu(_conv(sd, e + ".convc2", cor))
    flo = F.relu(_conv(sd, e + ".convf1", flow, padding=3))
    flo = F.relu(_conv(sd, e + ".convf2", flo))
    out = F.relu(_conv(sd, e + ".conv", torch.cat([cor, flo], dim=1)))
    x = torch.cat([inp, out, flow], dim=1)
    g = p + ".gru"
    for sfx, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([net, x], dim=1)
        z = torch.sigmoid(_conv(sd, g + ".convz" + sfx, hx, padding=pad))
        r = torch.sigmoid(_conv(sd, g + ".convr" + sfx, hx, padding=pad))
        q = torch.tanh(_conv(sd, g + ".convq" + sfx, torch.cat([r * net, x], dim=1), padding=pad))
        net = (1 - z) * net + z * q
    delta = _conv(sd, p + ".flow_head.conv2", F.relu(_conv(sd, p + ".flow_head.conv1", net)))
    mask = 0.25 * _conv(sd, p + ".mask.2", F.relu(_conv(sd, p + ".mask.0", net)), padding=0)
    return net, mask, delta


def raft_upsample_flow(flow, mask):
    """RAFT_SR.upsample_flow, raft_arch.py:720-731 (convex combination, 8x)"""
    N, _, H, W = flow.shape
    mask = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(N, 2, 9, 1, 1, H, W)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, 2, 8 * H, 8 * W)


def raft_forward(sd, ref, sup, iters=10, prefix=""):
    """RAFT_SR.forward / process (model='normal'), raft_arch.py:733-808: flow from `ref` to `sup`, (N,2,H,W) pixels."""
    sd = _strip(sd, prefix)
    ht, wd = ref.shape[-2:]
    pad_ht = (((ht // 8) + 1) * 8 - ht) % 8
    pad_wd = (((wd // 8) + 1) * 8 - wd) % 8
    pad = [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]      # InputPadder 'sintel', :17-27
    im1, im2 = F.pad(ref, pad, mode="replicate"), F.pad(sup, pad, mode="replicate")
    fm = raft_encoder(sd, "fnet", torch.cat([im1, im2], 0), "instance")
    fmap1, fmap2 = torch.split(fm.float(), [im1.shape[0], im1.shape[0]], dim=0)
    pyr = raft_corr_pyramid(fmap1, fmap2)
    cnet = raft_encoder(sd, "cnet", im1, "batch")
    net, inp = torch.split(cnet, [128, 128], dim=1)
    net, inp = torch.tanh(net), torch.relu(inp)
    N, _, H, W = im1.shape
    ys, xs = torch.meshgrid(torch.arange(H // 8), torch.arange(W // 8), indexing="ij")
    coords0 = torch.stack([xs, ys], dim=0).float()[None].repeat(N, 1, 1, 1).to(ref.device)
    coords1 = coords0.clone()
    flow_up = None
    for _ in range(iters):
        corr = raft_corr_lookup(pyr, coords1)
        net, mask, delta = _raft_update(sd, net, inp, corr, coords1 - coords0)
        coords1 = coords1 + delta
        flow_up = raft_upsample_flow(coords1 - coords0, mask)
    h2, w2 = flow_up.shape[-2:]
    return flow_up[..., pad[2]:h2 - pad[3], pad[0]:w2 - pad[1]]


def compute_flow(sd, lrs, prefix="flownet_model."):
    """LatentDiffusionVSRTextWT.compute_flow, ddpm.py:3404-3429"""
    n, t, c, h, w = lrs.shape
    l1, l2 = lrs[:, :-1].reshape(-1, c, h, w), lrs[:, 1:].reshape(-1, c, h, w)
    flows_backward = raft_forward(sd, l1, l2, prefix=prefix).view(n, t - 1, 2, h, w)
    flows_forward = raft_forward(sd, l2, l1, prefix=prefix).view(n, t - 1, 2, h, w)
    return flows_forward, flows_backward
