"""Host side of the conditional UNet and the struct-cond encoder: same constructor arguments, ``forward`` signatures
and ``state_dict`` key names as the reference classes, every tensor op dispatched to libmgld.so.

  InflatedUNetModelDualcondV2   ldm/modules/diffusionmodules/openaimodel.py:1903-2313
  InflatedEncoderUNetModelWT    ldm/modules/diffusionmodules/openaimodel.py:2316-2525

Activations live as fp16 NHWC ([T,H,W,C] == token layout [T,HW,C]); weights are packed once in ``load_state_dict``:
conv kernels tap-major/K-major, q|k|v concatenated, GEGLU value/gate and SPADE gamma/beta interleaved per 64 rows so
the GEMM epilogue sees both halves of a pair in one accumulator tile.  What is fused where:
  * conv bias + timestep-embedding add           -> conv_gemm bias vector (one GEMV per forward for all ResBlocks)
  * SPADE  GN(h)*(1+gamma)+beta + skip            -> epilogue of the gamma/beta conv
  * attention out-proj / FF2 / proj_out residuals -> epilogue (res, beta=1)
  * GEGLU                                         -> epilogue of the FF1 GEMM
  * temporal alpha-blend                          -> epilogue (alpha, 1-alpha, res=x)
  * the 16 cross-attention K/V projections of the constant text context -> ONE GEMM, cached per context tensor
"""
import torch

from . import ops as _cuda_ops
from .ops import (ACT_NONE, ACT_RELU, EPI_GEGLU, EPI_LINEAR, EPI_SPADE, TAPS_1, TAPS_3X3, TAPS_T3, interleave_pair,
                  pack_conv_weight, pack_temporal_weight)


def as_nhwc_f16(x, ops):
    """Accept the reference's NCHW fp32 tensors or our zero-copy NHWC fp16 (NCHW-shaped, channels-last strides)."""
    if x.dtype == torch.float16 and x.dim() == 4 and x.permute(0, 2, 3, 1).is_contiguous():
        return x.permute(0, 2, 3, 1)
    return ops.nchw_to_nhwc(x.float())


def nchw_view(x_nhwc):
    """NHWC fp16 storage presented with the reference's (N,C,H,W) shape (no copy)."""
    return x_nhwc.permute(0, 3, 1, 2)


class _Packed:
    """Parameter store: packs reference tensors into kernel layouts on the target device."""

    def __init__(self, sd, device):
        self.sd, self.dev = sd, device
        self.used = set()

    def has(self, k):
        return k in self.sd

    def raw(self, k):
        self.used.add(k)
        return self.sd[k]

    def f32(self, k):
        return self.raw(k).detach().to(self.dev, torch.float32).contiguous()

    def f16(self, k):
        return self.raw(k).detach().to(self.dev, torch.float16).contiguous()

    def conv(self, p):
        """-> (packed fp16 [Cout, taps*Cin], bias fp32)"""
        w = self.raw(p + ".weight").detach().float()
        if w.dim() == 3:      # conv1d k=1
            w = w[:, :, :, None]
        elif w.dim() == 2:    # linear
            w = w[:, :, None, None]
        b = self.f32(p + ".bias") if self.has(p + ".bias") else None
        return pack_conv_weight(w).to(self.dev), b

    def norm(self, p):
        return self.f32(p + ".weight"), self.f32(p + ".bias")


# Fusing the GroupNorm statistics of a conv's output into its epilogue (mgld_conv_gemm `stats_out`) removes one read
# pass + one launch per normalisation, but measured on B200 (profiles/r01_dev_run7_*.log) the extra epilogue work costs
# more than the standalone gn_stats kernels it replaces (UNet tile-step: conv_gemm +2.3 ms vs gn_stats -1.3 ms; VAE
# decode +14 ms), because the epilogue is on the critical path of the short-K GEMMs.  Off by default; kept as an option.
FUSE_GN_STATS = False


class StatsPool:
    """fp64 (sum, sumsq) slots for GroupNorm statistics that producers accumulate in their GEMM epilogues.
    One allocation + ONE memset per forward instead of a zero-fill launch per normalisation."""

    def __init__(self, slots=256, groups=32):
        self.slots, self.groups, self.buf, self.i = slots, groups, None, 0

    def reset(self, T, device):
        if not FUSE_GN_STATS:
            return
        if self.buf is None or self.buf.shape[1] != T or self.buf.device != torch.device(device):
            self.buf = torch.zeros(self.slots, T, self.groups, 2, 2, device=device, dtype=torch.float64)
        else:
            self.buf.zero_()
        self.i = 0

    def next(self):
        if not FUSE_GN_STATS:
            return None
        assert self.buf is not None and self.i < self.slots, "StatsPool exhausted / not reset"
        self.i += 1
        return self.buf[self.i - 1]


def _tag(t, sums):
    """remember the fused GroupNorm statistics of a producer's output on the tensor object"""
    if sums is not None:
        t._gn_sums = sums
    return t


def _gn_silu(ops, x, gb, eps, silu=True, x2=None):
    sums = getattr(x, "_gn_sums", None) if x2 is None else None
    if sums is not None:   # statistics came with the producer (FUSE_GN_STATS)
        return ops.gn_apply(x, sums, eps, gb[0], gb[1], silu, x2=x2)
    return ops.group_norm(x, gb[0], gb[1], eps, silu, x2=x2)


class _ResBlock:
    """ResBlock (openaimodel.py:233-360) and ResBlockDual (:362-482).  `dual` adds the SPADE tail."""

    def __init__(self, P, p, cin, cout, dual, emb_slices, spade_shared=None, level=0):
        self.cin, self.cout, self.dual = cin, cout, dual
        self.n1 = P.norm(p + ".in_layers.0")
        self.w1, b1 = P.conv(p + ".in_layers.2")
        self.n2 = P.norm(p + ".out_layers.0")
        self.w2, self.b2 = P.conv(p + ".out_layers.3")
        # h = conv1(.) + b1 + Linear(silu(emb)):  the GEMV of all blocks is batched by the owner (emb_slices)
        self.emb_idx = emb_slices.add(P.f16(p + ".emb_layers.1.weight"), P.f32(p + ".emb_layers.1.bias") + b1)
        self.skip = P.conv(p + ".skip_connection") if P.has(p + ".skip_connection.weight") else None
        if dual:
            s = p + ".spade"
            self.sn = P.norm(s + ".param_free_norm")
            self.ws, self.bs = P.conv(s + ".mlp_shared.0")
            self.sp_level, self.sp_slot = level, (spade_shared.add(level, self.ws, self.bs) if spade_shared is not None
                                                  else None)
            wg, bg = P.conv(s + ".mlp_gamma")
            wb, bb = P.conv(s + ".mlp_beta")
            self.wgb, self.bgb = interleave_pair(wg, wb), interleave_pair(bg, bb)

    def __call__(self, ops, x, emb_bias, seg=None, x2=None, pool=None, actv=None):
        """seg: the struct-cond map of this resolution; actv: ReLU(mlp_shared(seg)) if the owner already computed it"""
        a1 = _gn_silu(ops, x, self.n1, 1e-5, True, x2)
        T, HW = a1.shape[0], a1.shape[1] * a1.shape[2]
        s1 = pool.next() if pool is not None else None
        if s1 is None:
            s1 = ops.conv_stats_slot(T, HW, a1.device)      # statistics of h from the conv's own epilogue (long-K)
        h = _tag(ops.conv_gemm(a1, self.w1, taps=TAPS_3X3, bias=emb_bias[self.emb_idx], stats_out=s1), s1)
        a2 = _gn_silu(ops, h, self.n2, 1e-5, True)
        if self.skip is not None:
            sk = ops.conv_gemm(x, self.skip[0], taps=TAPS_1, a2=x2, bias=self.skip[1])
        else:
            assert x2 is None
            sk = x
        so = pool.next() if pool is not None else None
        if not self.dual:
            return _tag(ops.conv_gemm(a2, self.w2, taps=TAPS_3X3, bias=self.b2, res=sk, beta=1.0, stats_out=so), so)
        s2 = pool.next() if pool is not None else None
        if s2 is None:
            s2 = ops.conv_stats_slot(T, HW, a2.device)
        h2 = ops.conv_gemm(a2, self.w2, taps=TAPS_3X3, bias=self.b2, stats_out=s2)
        T, H, W, C = h2.shape
        st = ops.gn_finalize(s2, H * W, C, 1e-5) if s2 is not None else \
            ops.group_norm(h2, None, None, 1e-5, False, want_out=False, want_stats=True)
        if actv is None:
            actv = ops.conv_gemm(seg, self.ws, taps=TAPS_3X3, bias=self.bs, act=ACT_RELU)
        return _tag(ops.conv_gemm(actv, self.wgb, taps=TAPS_3X3, bias=self.bgb, epilogue=EPI_SPADE, h=h2, gn_stats=st,
                                  gn_weight=self.sn[0], gn_bias=self.sn[1], groups=32, res=sk, beta=1.0, stats_out=so), so)


class _EmbSlices:
    """Collects the emb_layers Linear(1280->Cout) of every ResBlock into one [sum Cout, E] GEMV."""

    def __init__(self):
        self.ws, self.bs, self.off = [], [], [0]

    def add(self, w, b):
        self.ws.append(w)
        self.bs.append(b)
        self.off.append(self.off[-1] + w.shape[0])
        return len(self.ws) - 1

    def finish(self):
        self.w = torch.cat(self.ws, 0).contiguous()
        self.b = torch.cat(self.bs, 0).contiguous()
        self.ws = self.bs = None

    def run(self, ops, emb):
        y = ops.gemv(emb, self.w, bias=self.b, silu_in=True)
        return [y[self.off[i]:self.off[i + 1]] for i in range(len(self.off) - 1)]


class _SpadeShared:
    """SPADE's first conv, ReLU(conv3x3(segmap 256 -> 128)) (spade.py:83-85), only depends on the struct-cond map of its
    resolution, not on the UNet activations: the 5-7 ResBlocks of one resolution level share ONE GEMM with their weights
    concatenated along N (a 128-wide GEMM runs at a third of the rate of a 640-wide one); each block reads its slice."""

    def __init__(self):
        self.groups = {}

    def add(self, level, w, b):
        g = self.groups.setdefault(level, [])
        g.append((w, b))
        return len(g) - 1

    def finish(self):
        self.w = {lv: torch.cat([w for w, _ in g], 0).contiguous() for lv, g in self.groups.items()}
        self.b = {lv: torch.cat([b for _, b in g], 0).contiguous() for lv, g in self.groups.items()}
        self.nh = {lv: g[0][0].shape[0] for lv, g in self.groups.items()}
        self.groups = None

    def run(self, ops, seg, width0):
        """seg: {width: NHWC map}; -> {level: [T,h,w,n_blocks*nhidden]}"""
        return {lv: ops.conv_gemm(seg[width0 >> lv], w, taps=TAPS_3X3, bias=self.b[lv], act=ACT_RELU)
                for lv, w in self.w.items()}

    def slice(self, all_actv, level, slot):
        nh = self.nh[level]
        return all_actv[level][..., slot * nh:(slot + 1) * nh]


class _KVCache:
    """Cross-attention K/V of the (constant) text context for all transformer blocks: one [ctx_len, sum 2C] GEMM."""

    def __init__(self):
        self.ws, self.off = [], [0]
        self.key, self.kv = None, None

    def add(self, wk, wv):
        self.ws += [wk, wv]
        self.off.append(self.off[-1] + wk.shape[0] + wv.shape[0])
        return len(self.off) - 2

    def finish(self):
        self.w = torch.cat(self.ws, 0).contiguous() if self.ws else None
        self.ws = None

    def get(self, ops, context):
        # keyed on the tensor OBJECT (kept alive here, so its address cannot be recycled by another tensor) + its version
        if self.key is None or self.key[0] is not context or self.key[1] != context._version:
            assert context.shape[0] == 1, "one text context per clip (SURVEY.md D4/D6)"
            ctx = context[0].to(self.w.device, torch.float16).contiguous()
            self.kv = ops.conv_gemm(ctx, self.w)          # [ctx_len, sum 2C]
            self.key = (context, context._version)
        return self.kv


class _SpatialTransformer:
    """SpatialTransformerV2 (use_linear, depth 1), attention.py:484-546 + BasicTransformerBlockV2 :406-435."""

    def __init__(self, P, p, C, heads, kvc):
        self.C, self.heads = C, heads
        self.norm = P.norm(p + ".norm")
        self.win, self.bin = P.conv(p + ".proj_in")
        self.wout, self.bout = P.conv(p + ".proj_out")
        b = p + ".transformer_blocks.0"
        self.ln = [P.norm(f"{b}.norm{i}") for i in (1, 2, 3)]
        self.wqkv = torch.cat([P.f16(f"{b}.attn1.to_{n}.weight") for n in "qkv"], 0).contiguous()
        self.wo1, self.bo1 = P.conv(b + ".attn1.to_out.0")
        self.wq2 = P.f16(b + ".attn2.to_q.weight")
        self.kv_idx = kvc.add(P.f16(b + ".attn2.to_k.weight"), P.f16(b + ".attn2.to_v.weight"))
        self.wo2, self.bo2 = P.conv(b + ".attn2.to_out.0")
        wf, bf = P.f16(b + ".ff.net.0.proj.weight"), P.f32(b + ".ff.net.0.proj.bias")
        n = wf.shape[0] // 2
        self.wff1, self.bff1 = interleave_pair(wf[:n], wf[n:]), interleave_pair(bf[:n], bf[n:])
        self.wff2, self.bff2 = P.conv(b + ".ff.net.2")

    def __call__(self, ops, x, kv, kvc, pool=None):
        T, H, W, C = x.shape
        N, heads = H * W, self.heads
        dh = C // heads
        scale = dh ** -0.5
        xn = _gn_silu(ops, x, self.norm, 1e-6, silu=False)
        t = ops.conv_gemm(xn.reshape(T * N, C), self.win, bias=self.bin)
        # self-attention
        qkv = ops.conv_gemm(ops.layernorm(t, *self.ln[0]), self.wqkv)
        a = ops.attention(qkv, qkv, qkv, batch=T, heads=heads, head_dim=dh, nq=N, nkv=N, scale=scale, q_col0=0,
                          k_col0=C, v_col0=2 * C)
        t = ops.conv_gemm(a, self.wo1, bias=self.bo1, res=t, beta=1.0)
        # cross-attention to the text context (K/V broadcast to every frame, attention.py:336-337)
        q = ops.conv_gemm(ops.layernorm(t, *self.ln[1]), self.wq2)
        off = kvc.off[self.kv_idx]
        a = ops.attention(q, kv, kv, batch=T, heads=heads, head_dim=dh, nq=N, nkv=kv.shape[0], scale=scale,
                          k_col0=off, v_col0=off + C, kv_batched=False)
        t = ops.conv_gemm(a, self.wo2, bias=self.bo2, res=t, beta=1.0)
        # GEGLU feed-forward
        g = ops.conv_gemm(ops.layernorm(t, *self.ln[2]), self.wff1, bias=self.bff1, epilogue=EPI_GEGLU)
        t = ops.conv_gemm(g, self.wff2, bias=self.bff2, res=t, beta=1.0)
        so = pool.next() if pool is not None else None
        out = ops.conv_gemm(t.reshape(T, N, C), self.wout, bias=self.bout, res=x.reshape(T, N, C), beta=1.0, stats_out=so)
        return _tag(out.reshape(T, H, W, C), so)


class _TemporalConv:
    """SpatialTemporalConv, util.py:291-310: alpha * conv3d_(3,1,1)(x) + (1-alpha) * x.
    x holds `(b t)` frames (b clips of `nf` frames): the temporal zero padding applies per clip, so each clip is one
    GEMM whose TMA box loads run out of bounds at ITS first / last frame."""

    def __init__(self, P, p, nf):
        self.w = pack_temporal_weight(P.raw(p + ".temporal_conv.weight").detach().float()).to(P.dev)
        self.b = P.f32(p + ".temporal_conv.bias")
        self.alpha = float(P.raw(p + ".temporal_alpha").detach().float().reshape(-1)[0])
        self.nf = nf

    def __call__(self, ops, x, pool=None):
        so = pool.next() if pool is not None else None
        nf = self.nf
        if x.shape[0] <= nf:
            return _tag(ops.conv_gemm(x, self.w, taps=TAPS_T3, bias=self.b, alpha=self.alpha, beta=1.0 - self.alpha,
                                      res=x, stats_out=so), so)
        assert x.shape[0] % nf == 0 and so is None, "frames must be whole clips of num_frames"
        out = torch.empty_like(x)
        for k in range(0, x.shape[0], nf):
            ops.conv_gemm(x[k:k + nf], self.w, taps=TAPS_T3, bias=self.b, alpha=self.alpha, beta=1.0 - self.alpha,
                          res=x[k:k + nf], out=out[k:k + nf])
        return out


class _TemporalAttention:
    """TemporalAttention, attention.py:124-143: `(b t) c h w -> (b h w) t c`, attention along t within each clip."""

    def __init__(self, P, p, C, heads, nf):
        self.heads, self.nf = heads, nf
        self.ln = P.norm(p + ".norm")
        a = p + ".temporal_attn"
        self.wqkv = torch.cat([P.f16(f"{a}.to_{n}.weight") for n in "qkv"], 0).contiguous()
        self.wo, self.bo = P.conv(a + ".to_out.0")
        self.alpha = float(P.raw(p + ".temporal_alpha").detach().float().reshape(-1)[0])

    def __call__(self, ops, x, pool=None):
        T, H, W, C = x.shape
        nf = self.nf
        n = ops.layernorm(x.reshape(T * H * W, C), *self.ln)
        qkv = ops.conv_gemm(n, self.wqkv).reshape(T, H * W, 3 * C)
        scale = (C // self.heads) ** -0.5
        if T <= nf:
            a = ops.temporal_attention(qkv, self.heads, scale)
        else:
            assert T % nf == 0, "frames must be whole clips of num_frames"
            a = torch.cat([ops.temporal_attention(qkv[k:k + nf], self.heads, scale) for k in range(0, T, nf)], 0)
        so = pool.next() if pool is not None else None
        out = ops.conv_gemm(a, self.wo, bias=self.bo, alpha=self.alpha, beta=1.0 - self.alpha,
                            res=x.reshape(T, H * W, C), stats_out=so)
        return _tag(out.reshape(T, H, W, C), so)


class _Downsample:
    def __init__(self, P, p):
        self.w, self.b = P.conv(p + ".op")

    def __call__(self, ops, x, pool=None):
        so = pool.next() if pool is not None else None
        return _tag(ops.conv_gemm(ops.im2col_s2(x, 1), self.w, taps=TAPS_1, bias=self.b, stats_out=so), so)


class _Upsample:
    def __init__(self, P, p):
        self.w, self.b = P.conv(p + ".conv")

    def __call__(self, ops, x, pool=None):
        so = pool.next() if pool is not None else None
        return _tag(ops.conv_gemm(ops.upsample2x(x), self.w, taps=TAPS_3X3, bias=self.b, stats_out=so), so)


class _TimeEmbed:
    def __init__(self, P, p, model_channels):
        self.mc = model_channels
        self.w0, self.b0 = P.f16(p + ".0.weight"), P.f32(p + ".0.bias")
        self.w2, self.b2 = P.f16(p + ".2.weight"), P.f32(p + ".2.bias")

    def __call__(self, ops, t_dev):
        te = ops.timestep_embedding(t_dev, self.mc)
        return ops.gemv(ops.gemv(te, self.w0, bias=self.b0, silu_out=True), self.w2, bias=self.b2)


def _t_scalar(timesteps, device):
    """timesteps: 1-element tensor (the reference feeds ts of shape (b,), b == 1: SURVEY.md D4) -> fp32 device scalar."""
    # the untiled sampler repeats the same value T times (ddpm.py:4533); all entries are equal by construction
    return timesteps.reshape(-1)[:1].to(device=device, dtype=torch.float32)


class _ModuleBase:
    """Minimal stand-in for the nn.Module surface the reference scripts touch."""

    def eval(self):
        return self

    def cuda(self):
        return self

    def to(self, *a, **k):
        return self

    def half(self):
        return self

    def parameters(self):
        return iter(())


class InflatedUNetModelDualcondV2(_ModuleBase):
    """SD-2.1 UNet with SPADE ResBlocks + temporal layers in the middle block (openaimodel.py:1903)."""

    def __init__(self, image_size=32, in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2,
                 attention_resolutions=(4, 2, 1), dropout=0, channel_mult=(1, 2, 4, 4), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=64,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, use_spatial_transformer=True, transformer_depth=1, context_dim=1024,
                 n_embed=None, legacy=False, disable_self_attentions=None, num_attention_blocks=None,
                 disable_middle_self_attn=False, use_linear_in_transformer=True, semb_channels=256, num_frames=5,
                 ops=None, **ignored):
        if not (use_spatial_transformer and use_linear_in_transformer and transformer_depth == 1 and not legacy
                and not use_scale_shift_norm and not resblock_updown and num_classes is None and dims == 2):
            raise NotImplementedError("only the configuration shipped in configs/mgldvsr/*.yaml is implemented")
        if num_head_channels != 64:
            raise NotImplementedError("attention head_dim must be 64 (tcgen05 attention kernel)")
        self.cfg = dict(in_channels=in_channels, model_channels=model_channels, out_channels=out_channels,
                        num_res_blocks=num_res_blocks, attention_resolutions=list(attention_resolutions),
                        channel_mult=list(channel_mult), num_head_channels=num_head_channels, context_dim=context_dim,
                        semb_channels=semb_channels, num_frames=num_frames)
        self.num_frames = num_frames
        self.ops = ops or _cuda_ops
        self.pool = StatsPool()
        self.loaded = False

    # ---- structure (mirrors openaimodel.py:2036-2257) -------------------------------------------------------------
    def layout(self):
        c = self.cfg
        mc, mult, nrb, attn_res, nhc = (c["model_channels"], c["channel_mult"], c["num_res_blocks"],
                                        c["attention_resolutions"], c["num_head_channels"])
        inp, chans, ch, ds = [[("conv_in",)]], [mc], mc, 1
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
                layers = [("res", ch + ich, mc * m, ch)]      # 4th entry: channels coming from h (rest = skip)
                ch = mc * m
                if ds in attn_res:
                    layers.append(("st", ch, ch // nhc))
                if level and i == nrb:
                    layers.append(("up", ch))
                    ds //= 2
                out.append(layers)
        return inp, mid, out

    def expected_shapes(self):
        """state_dict manifest {key: shape} — lets tests build weights without the reference."""
        c = self.cfg
        mc, E, S, ctx = c["model_channels"], c["model_channels"] * 4, c["semb_channels"], c["context_dim"]
        sh = {"time_embed.0.weight": (E, mc), "time_embed.0.bias": (E,), "time_embed.2.weight": (E, E),
              "time_embed.2.bias": (E,)}

        def conv(p, co, ci, k=3):
            sh[p + ".weight"], sh[p + ".bias"] = (co, ci, k, k), (co,)

        def norm(p, cch):
            sh[p + ".weight"], sh[p + ".bias"] = (cch,), (cch,)

        def lin(p, co, ci, bias=True):
            sh[p + ".weight"] = (co, ci)
            if bias:
                sh[p + ".bias"] = (co,)

        def add(p, l):
            if l[0] == "conv_in":
                conv(p, mc, c["in_channels"])
            elif l[0] == "res":
                ci, co = l[1], l[2]
                norm(p + ".in_layers.0", ci); conv(p + ".in_layers.2", co, ci)
                lin(p + ".emb_layers.1", co, E)
                norm(p + ".out_layers.0", co); conv(p + ".out_layers.3", co, co)
                norm(p + ".spade.param_free_norm", co); conv(p + ".spade.mlp_shared.0", 128, S)
                conv(p + ".spade.mlp_gamma", co, 128); conv(p + ".spade.mlp_beta", co, 128)
                if ci != co:
                    conv(p + ".skip_connection", co, ci, 1)
            elif l[0] == "st":
                C = l[1]
                norm(p + ".norm", C); lin(p + ".proj_in", C, C); lin(p + ".proj_out", C, C)
                b = p + ".transformer_blocks.0"
                for i in (1, 2, 3):
                    norm(f"{b}.norm{i}", C)
                for n in "qkv":
                    lin(f"{b}.attn1.to_{n}", C, C, False)
                lin(b + ".attn1.to_out.0", C, C)
                lin(b + ".attn2.to_q", C, C, False); lin(b + ".attn2.to_k", C, ctx, False); lin(b + ".attn2.to_v", C, ctx, False)
                lin(b + ".attn2.to_out.0", C, C)
                lin(b + ".ff.net.0.proj", 8 * C, C); lin(b + ".ff.net.2", C, 4 * C)
            elif l[0] == "stconv":
                C = l[1]
                sh[p + ".temporal_conv.weight"], sh[p + ".temporal_conv.bias"] = (C, C, 3, 1, 1), (C,)
                sh[p + ".temporal_alpha"] = (1,)
            elif l[0] == "tattn":
                C = l[1]
                norm(p + ".norm", C)
                for n in "qkv":
                    lin(f"{p}.temporal_attn.to_{n}", C, C, False)
                lin(p + ".temporal_attn.to_out.0", C, C)
                sh[p + ".temporal_alpha"] = (1,)
            elif l[0] == "down":
                conv(p + ".op", l[1], l[1])
            elif l[0] == "up":
                conv(p + ".conv", l[1], l[1])

        inp, mid, out = self.layout()
        for i, layers in enumerate(inp):
            for j, l in enumerate(layers):
                add(f"input_blocks.{i}.{j}", l)
        for j, l in enumerate(mid):
            add(f"middle_block.{j}", l)
        for i, layers in enumerate(out):
            for j, l in enumerate(layers):
                add(f"output_blocks.{i}.{j}", l)
        norm("out.0", mc)
        conv("out.2", c["out_channels"], mc)
        return sh

    # ---- weights -----------------------------------------------------------------------------------------------------
    def load_state_dict(self, sd, strict=True, device="cuda"):
        P = _Packed(sd, torch.device(device))
        self.emb = _EmbSlices()
        self.kvc = _KVCache()
        self.time_embed = _TimeEmbed(P, "time_embed", self.cfg["model_channels"])

        self.spade = _SpadeShared()
        level = [0]                                    # resolution level of the layer being built (0 = input size)

        def build(p, l):
            if l[0] == "conv_in":
                return ("conv_in", (P.f32(p + ".weight"), P.f32(p + ".bias")))
            if l[0] == "res":
                return ("res", _ResBlock(P, p, l[1], l[2], True, self.emb, self.spade, level[0]),
                        l[3] if len(l) > 3 else None)
            if l[0] == "st":
                return ("st", _SpatialTransformer(P, p, l[1], l[2], self.kvc))
            if l[0] == "stconv":
                return ("stconv", _TemporalConv(P, p, self.num_frames))
            if l[0] == "tattn":
                return ("tattn", _TemporalAttention(P, p, l[1], l[2], self.num_frames))
            if l[0] == "down":
                level[0] += 1
                return ("down", _Downsample(P, p))
            if l[0] == "up":
                level[0] -= 1
                return ("up", _Upsample(P, p))
            raise ValueError(l)

        inp, mid, out = self.layout()
        self.input_blocks = [[build(f"input_blocks.{i}.{j}", l) for j, l in enumerate(ls)] for i, ls in enumerate(inp)]
        self.middle_block = [build(f"middle_block.{j}", l) for j, l in enumerate(mid)]
        self.output_blocks = [[build(f"output_blocks.{i}.{j}", l) for j, l in enumerate(ls)] for i, ls in enumerate(out)]
        self.out_norm = P.norm("out.0")
        self.out_w = pack_conv_weight(P.raw("out.2.weight").detach().float()).to(P.dev)
        self.out_b = P.f32("out.2.bias")
        self.emb.finish()
        self.kvc.finish()
        self.spade.finish()
        assert level[0] == 0
        self.loaded = True
        missing = [k for k in self.expected_shapes() if k not in sd]
        unexpected = [k for k in sd if k not in P.used]
        if strict and (missing or unexpected):
            raise KeyError(f"state_dict mismatch: missing {missing[:5]}..., unexpected {unexpected[:5]}...")
        return missing, unexpected

    # ---- forward (openaimodel.py:2281-2313) ---------------------------------------------------------------------------
    def _run(self, layers, h, emb_bias, kv, seg, h2=None):
        ops = self.ops
        for l in layers:
            kind, mod = l[0], l[1]
            if kind == "conv_in":
                h = ops.conv_small_cin(h, mod[0], mod[1])
            elif kind == "res":
                h = mod(ops, h, emb_bias, seg[h.shape[2]], x2=h2, pool=self.pool,
                        actv=self.spade.slice(self._actv, mod.sp_level, mod.sp_slot))
                h2 = None
            elif kind == "st":
                h = mod(ops, h, kv, self.kvc, pool=self.pool)
            else:
                h = mod(ops, h, pool=self.pool)
        return h

    def forward(self, x, timesteps=None, context=None, struct_cond=None, y=None, **kwargs):
        assert self.loaded, "load_state_dict() first"
        assert y is None, "class-conditional models are not supported"
        ops = self.ops
        self.pool.reset(x.shape[0], x.device)
        ops.stats_pool_reset()
        seg = {int(k): as_nhwc_f16(v, ops) for k, v in struct_cond.items()}
        emb = self.time_embed(ops, _t_scalar(timesteps, x.device))
        emb_bias = self.emb.run(ops, emb)
        kv = self.kvc.get(ops, context)
        self._actv = self.spade.run(ops, seg, x.shape[3])
        hs, h = [], x.float().contiguous()
        for layers in self.input_blocks:
            h = self._run(layers, h, emb_bias, kv, seg)
            hs.append(h)
        h = self._run(self.middle_block, h, emb_bias, kv, seg)
        for layers in self.output_blocks:
            h = self._run(layers, h, emb_bias, kv, seg, h2=hs.pop())   # th.cat([h, hs.pop()], dim=1) fused as 2 sources
        self._actv = None
        a = _gn_silu(ops, h, self.out_norm, 1e-5, True)
        return ops.conv3x3_small_cout(a, self.out_w, self.out_b)       # (T, out_ch, H, W) fp32

    __call__ = forward


class _AttentionBlock:
    """AttentionBlock + QKVAttentionLegacy (openaimodel.py:485-590); qkv columns are (head, {q,k,v}, ch)."""

    def __init__(self, P, p, C, heads):
        self.C, self.heads = C, heads
        self.norm = P.norm(p + ".norm")
        self.wqkv, self.bqkv = P.conv(p + ".qkv")
        self.wo, self.bo = P.conv(p + ".proj_out")

    def __call__(self, ops, x, pool=None):
        T, H, W, C = x.shape
        N, ch = H * W, C // self.heads
        xn = _gn_silu(ops, x, self.norm, 1e-5, silu=False)
        qkv = ops.conv_gemm(xn.reshape(T * N, C), self.wqkv, bias=self.bqkv)
        a = ops.attention(qkv, qkv, qkv, batch=T, heads=self.heads, head_dim=ch, nq=N, nkv=N, scale=ch ** -0.5,
                          q_col0=0, k_col0=ch, v_col0=2 * ch, q_head_stride=3 * ch, k_head_stride=3 * ch,
                          v_head_stride=3 * ch)
        so = pool.next() if pool is not None else None
        out = ops.conv_gemm(a.reshape(T, N, C), self.wo, bias=self.bo, res=x.reshape(T, N, C), beta=1.0, stats_out=so)
        return _tag(out.reshape(T, H, W, C), so)


class InflatedEncoderUNetModelWT(_ModuleBase):
    """Time-aware struct-cond encoder of the LR latent (openaimodel.py:2316-2525)."""

    def __init__(self, image_size=96, in_channels=4, model_channels=256, out_channels=256, num_res_blocks=2,
                 attention_resolutions=(4, 2, 1), dropout=0, channel_mult=(1, 1, 2, 2), conv_resample=True, dims=2,
                 use_checkpoint=False, use_fp16=False, num_heads=4, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False, num_frames=5,
                 ops=None, **ignored):
        if use_scale_shift_norm or resblock_updown or use_new_attention_order or num_head_channels != -1 or dims != 2:
            raise NotImplementedError("only the configuration shipped in configs/mgldvsr/*.yaml is implemented")
        self.cfg = dict(in_channels=in_channels, model_channels=model_channels, out_channels=out_channels,
                        num_res_blocks=num_res_blocks, attention_resolutions=list(attention_resolutions),
                        channel_mult=list(channel_mult), num_heads=num_heads, num_frames=num_frames)
        self.ops = ops or _cuda_ops
        self.pool = StatsPool()
        self.loaded = False

    def layout(self):
        c = self.cfg
        mc, mult, nrb, attn_res = c["model_channels"], c["channel_mult"], c["num_res_blocks"], c["attention_resolutions"]
        blocks, chans, ch, ds = [[("conv_in",)]], [], mc, 1
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

    def expected_shapes(self):
        c = self.cfg
        mc, E = c["model_channels"], c["model_channels"] * 4
        sh = {"time_embed.0.weight": (E, mc), "time_embed.0.bias": (E,), "time_embed.2.weight": (E, E),
              "time_embed.2.bias": (E,)}

        def res(p, ci, co):
            sh[p + ".in_layers.0.weight"] = sh[p + ".in_layers.0.bias"] = (ci,)
            sh[p + ".in_layers.2.weight"], sh[p + ".in_layers.2.bias"] = (co, ci, 3, 3), (co,)
            sh[p + ".emb_layers.1.weight"], sh[p + ".emb_layers.1.bias"] = (co, E), (co,)
            sh[p + ".out_layers.0.weight"] = sh[p + ".out_layers.0.bias"] = (co,)
            sh[p + ".out_layers.3.weight"], sh[p + ".out_layers.3.bias"] = (co, co, 3, 3), (co,)
            if ci != co:
                sh[p + ".skip_connection.weight"], sh[p + ".skip_connection.bias"] = (co, ci, 1, 1), (co,)

        def attn(p, C):
            sh[p + ".norm.weight"] = sh[p + ".norm.bias"] = (C,)
            sh[p + ".qkv.weight"], sh[p + ".qkv.bias"] = (3 * C, C, 1), (3 * C,)
            sh[p + ".proj_out.weight"], sh[p + ".proj_out.bias"] = (C, C, 1), (C,)

        blocks, ch, chans = self.layout()
        for i, layers in enumerate(blocks):
            for j, l in enumerate(layers):
                p = f"input_blocks.{i}.{j}"
                if l[0] == "conv_in":
                    sh[p + ".weight"], sh[p + ".bias"] = (mc, c["in_channels"], 3, 3), (mc,)
                elif l[0] == "res":
                    res(p, l[1], l[2])
                elif l[0] == "attn":
                    attn(p, l[1])
                elif l[0] == "down":
                    sh[p + ".op.weight"], sh[p + ".op.bias"] = (l[1], l[1], 3, 3), (l[1],)
        res("middle_block.0", ch, ch); attn("middle_block.1", ch); res("middle_block.2", ch, ch)
        for i, cc in enumerate(chans):
            res(f"fea_tran.{i}", cc, c["out_channels"])
        return sh

    def load_state_dict(self, sd, strict=True, device="cuda"):
        P = _Packed(sd, torch.device(device))
        self.emb = _EmbSlices()
        self.time_embed = _TimeEmbed(P, "time_embed", self.cfg["model_channels"])
        heads = self.cfg["num_heads"]
        blocks, ch, chans = self.layout()
        self.blocks = []
        for i, layers in enumerate(blocks):
            mods = []
            for j, l in enumerate(layers):
                p = f"input_blocks.{i}.{j}"
                if l[0] == "conv_in":
                    mods.append(("conv_in", (P.f32(p + ".weight"), P.f32(p + ".bias"))))
                elif l[0] == "res":
                    mods.append(("res", _ResBlock(P, p, l[1], l[2], False, self.emb)))
                elif l[0] == "attn":
                    mods.append(("attn", _AttentionBlock(P, p, l[1], heads)))
                elif l[0] == "down":
                    mods.append(("down", _Downsample(P, p)))
            self.blocks.append(mods)
        self.mid = [("res", _ResBlock(P, "middle_block.0", ch, ch, False, self.emb)),
                    ("attn", _AttentionBlock(P, "middle_block.1", ch, heads)),
                    ("res", _ResBlock(P, "middle_block.2", ch, ch, False, self.emb))]
        self.fea_tran = [_ResBlock(P, f"fea_tran.{i}", cc, self.cfg["out_channels"], False, self.emb)
                         for i, cc in enumerate(chans)]
        self.emb.finish()
        self.loaded = True
        missing = [k for k in self.expected_shapes() if k not in sd]
        unexpected = [k for k in sd if k not in P.used]
        if strict and (missing or unexpected):
            raise KeyError(f"state_dict mismatch: missing {missing[:5]}..., unexpected {unexpected[:5]}...")
        return missing, unexpected

    def _run(self, mods, h, emb_bias):
        ops = self.ops
        for kind, mod in mods:
            if kind == "conv_in":
                h = ops.conv_small_cin(h, mod[0], mod[1])
            elif kind == "res":
                h = mod(ops, h, emb_bias, pool=self.pool)
            else:
                h = mod(ops, h, pool=self.pool)
        return h

    def forward(self, x, timesteps):
        """-> {'64': (T,256,64,64), '32': ..., ...}: NCHW-shaped fp16 tensors with channels-last storage."""
        assert self.loaded, "load_state_dict() first"
        ops = self.ops
        self.pool.reset(x.shape[0], x.device)
        ops.stats_pool_reset()
        emb_bias = self.emb.run(ops, self.time_embed(ops, _t_scalar(timesteps, x.device)))
        results, h = [], x.float().contiguous()
        for mods in self.blocks:
            last = h
            h = self._run(mods, h, emb_bias)
            if last.dim() == 4 and last.dtype == torch.float16 and h.shape[2] != last.shape[2]:
                results.append(last)
        h = self._run(self.mid, h, emb_bias)
        results.append(h)
        assert len(results) == len(self.fea_tran)
        return {str(r.shape[2]): nchw_view(self.fea_tran[i](ops, r, emb_bias, pool=self.pool)) for i, r in enumerate(results)}

    __call__ = forward
