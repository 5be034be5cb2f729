"""Host side of the two KL autoencoders, with the reference's class names, constructor arguments, ``encode`` /
``decode`` signatures and state_dict keys:

  AutoencoderKL            ldm/models/autoencoder.py:299-360   (LR clip -> latent, ``first_stage_model``)
  VideoAutoencoderKLResi   ldm/models/autoencoder.py:1564-1700 (``vq_model``: Encoder with feature taps +
                           VideoDecoder_Mix = SD decoder + SpatialTemporalConv after every ResnetBlock + two
                           encoder-feature fusion blocks; ldm/modules/diffusionmodules/model.py:473-572, 926-1056)

All convolutions run through the tcgen05 implicit-GEMM kernel; the single-head head-dim-512 middle attention is one
fused q|k|v projection GEMM + the split-D flash kernel of csrc/attention_hd512.cu (head widths without a fused kernel
fall back to GEMM (fp32 scores) -> row softmax -> GEMM per L2-sized query panel, with V^T produced directly by an
operand-swapped GEMM and the V bias added after P V, exact because softmax rows sum to one).  The ResidualDenseBlock concat chain is a
single channel-slotted buffer: each growth conv writes its 32 channels into its own 64-wide slot, so no torch.cat and
no copy ever happens.
"""
import os

import torch

from . import ops as _cuda_ops
from .ops import ACT_LRELU02, TAPS_1, TAPS_3X3, pack_conv_weight
from .unet import StatsPool, _ModuleBase, _Packed, _TemporalConv, _Upsample, _gn_silu, _tag, as_nhwc_f16, nchw_view


class DiagonalGaussianDistribution:
    """ldm/modules/distributions/distributions.py:24-60 (the parts inference touches)."""

    def __init__(self, parameters, ops, deterministic=False):
        self.parameters, self.ops, self.deterministic = parameters, ops, deterministic
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)

    def sample(self):
        # same RNG stream as the reference: a CPU-generator draw moved to the device (distributions.py:36)
        noise = None if self.deterministic else torch.randn(self.mean.shape).to(device=self.parameters.device)
        return self.ops.gaussian_sample(self.parameters, noise, 1.0)

    def mode(self):
        return self.mean


class _ResnetBlock:
    """ResnetBlock (temb=None), model.py:124-183; also the fusion ResBlock, model.py:1312-1335 (skip key differs)."""

    def __init__(self, P, p, skip_key="nin_shortcut"):
        self.n1, self.n2 = P.norm(p + ".norm1"), P.norm(p + ".norm2")
        self.w1, self.b1 = P.conv(p + ".conv1")
        self.w2, self.b2 = P.conv(p + ".conv2")
        self.skip = P.conv(f"{p}.{skip_key}") if P.has(f"{p}.{skip_key}.weight") else None

    def __call__(self, ops, x, x2=None, out=None, pool=None):
        a1 = _gn_silu(ops, x, self.n1, 1e-6, True, x2)
        s1 = pool.next() if pool is not None else None
        if s1 is None:
            s1 = ops.conv_stats_slot(a1.shape[0], a1.shape[1] * a1.shape[2], a1.device)   # sums of h from conv1's epilogue
        h = _tag(ops.conv_gemm(a1, self.w1, taps=TAPS_3X3, bias=self.b1, stats_out=s1), s1)
        a2 = _gn_silu(ops, h, self.n2, 1e-6, True)
        if self.skip is not None:
            sk = ops.conv_gemm(x, self.skip[0], taps=TAPS_1, a2=x2, bias=self.skip[1])
        else:
            assert x2 is None
            sk = x
        so = pool.next() if (pool is not None and out is None) else None
        return _tag(ops.conv_gemm(a2, self.w2, taps=TAPS_3X3, bias=self.b2, res=sk, beta=1.0, out=out, stats_out=so), so)


class _AttnBlock:
    """MemoryEfficientAttnBlock, model.py:247-305: one head of width C, scale C^-0.5."""

    def __init__(self, P, p):
        self.norm = P.norm(p + ".norm")
        wq, bq = P.conv(p + ".q")
        wk, bk = P.conv(p + ".k")
        wv, bv = P.conv(p + ".v")
        C = wq.shape[0]
        # one [3C, C] weight: the fused path projects q | k | v with ONE GEMM; the panelled path uses its row blocks
        self.wqkv, self.bqkv = torch.cat([wq, wk, wv], 0).contiguous(), torch.cat([bq, bk, bv], 0).contiguous()
        self.wq, self.wk, self.wv = self.wqkv[:C], self.wqkv[C:2 * C], self.wqkv[2 * C:]
        self.bq, self.bk, self.bv = self.bqkv[:C], self.bqkv[C:2 * C], self.bqkv[2 * C:]
        self.wo, self.bo = P.conv(p + ".proj_out")

    # Bytes of fp32 scores per query panel.  Queries are processed in panels of `rows` x N scores so that the softmax and the
    # PV GEMM read the scores / probabilities from the 126 MB L2 instead of HBM.  Measured on one frame of a 960x960 tile
    # (N = 14400, ncu without cache flushing, profiles/r02_ncu_vae_attention_960*.csv): one panel per frame (the r01
    # behaviour) moves 2.74 GB of reads + 1.42 GB of writes through DRAM; 32 MB panels 0.65 GB + 1.27 GB (the write-back L2
    # still cleans every dirty score line to DRAM) — against 59 MB of Q + K + V + O.  64 MB keeps a 512x512 frame (N = 4096)
    # in one panel.  Only head widths without a fused kernel (see FUSED_HEAD_DIMS) still take this path.
    PANEL_BYTES = int(os.environ.get("MGLD_VAE_PANEL_MB", "64")) << 20
    # head widths with a fused attention kernel; MGLD_VAE_FUSED_ATTN=0 forces the panelled GEMM -> softmax -> GEMM path
    FUSED_HEAD_DIMS = (64, 128, 512) if os.environ.get("MGLD_VAE_FUSED_ATTN", "1") != "0" else ()

    def __call__(self, ops, x):
        T, H, W, C = x.shape
        N = H * W
        xn = _gn_silu(ops, x, self.norm, 1e-6, silu=False).reshape(T * N, C)
        if C in self.FUSED_HEAD_DIMS:
            # flash kernel (C = 512: csrc/attention_hd512.cu): no score matrix in memory
            qkv = ops.conv_gemm(xn, self.wqkv, bias=self.bqkv)
            o = ops.attention(qkv, qkv, qkv, batch=T, heads=1, head_dim=C, nq=N, nkv=N, scale=float(C) ** -0.5,
                              q_col0=0, k_col0=C, v_col0=2 * C)
            out = ops.conv_gemm(o, self.wo, bias=self.bo, res=x.reshape(T * N, C), beta=1.0)
            return out.reshape(T, H, W, C)
        q = ops.conv_gemm(xn, self.wq, bias=self.bq)
        k = ops.conv_gemm(xn, self.wk, bias=self.bk)
        o = torch.empty(T * N, C, device=x.device, dtype=torch.float16)
        panel = max(128, min(N, (self.PANEL_BYTES // (4 * N)) // 128 * 128))
        for t in range(T):
            vT = ops.conv_gemm(self.wv, xn[t * N:(t + 1) * N])           # V^T = Wv Xn^T : [C, N]
            kt = k[t * N:(t + 1) * N]
            for r0 in range(t * N, (t + 1) * N, panel):
                rows = slice(r0, min(r0 + panel, (t + 1) * N))
                s = ops.conv_gemm(q[rows], kt, out_f32=True)             # [panel, N] fp32 scores (L2-resident)
                p = ops.softmax_rows(s, float(C) ** -0.5)                # [panel, N] fp16
                ops.conv_gemm(p, vT, bias=self.bv, out=o[rows])          # P V + bv
        out = ops.conv_gemm(o, self.wo, bias=self.bo, res=x.reshape(T * N, C), beta=1.0)
        return out.reshape(T, H, W, C)


class _VaeDownsample:
    """Downsample, model.py:104-121: F.pad(x,(0,1,0,1)) + conv3x3 stride 2."""

    def __init__(self, P, p):
        self.w, self.b = P.conv(p + ".conv")

    def __call__(self, ops, x, pool=None):
        so = pool.next() if pool is not None else None
        return _tag(ops.conv_gemm(ops.im2col_s2(x, 0), self.w, taps=TAPS_1, bias=self.b, stats_out=so), so)


class _Encoder:
    """Encoder, model.py:473-572."""

    def __init__(self, P, p, dd):
        self.nres, self.nrb = len(dd["ch_mult"]), dd["num_res_blocks"]
        self.conv_in = (P.f32(p + "conv_in.weight"), P.f32(p + "conv_in.bias"))
        self.down = []
        for lvl in range(self.nres):
            blocks = [_ResnetBlock(P, f"{p}down.{lvl}.block.{b}") for b in range(self.nrb)]
            ds = _VaeDownsample(P, f"{p}down.{lvl}.downsample") if lvl != self.nres - 1 else None
            self.down.append((blocks, ds))
        self.mid1, self.attn, self.mid2 = (_ResnetBlock(P, p + "mid.block_1"), _AttnBlock(P, p + "mid.attn_1"),
                                           _ResnetBlock(P, p + "mid.block_2"))
        self.norm_out = P.norm(p + "norm_out")
        self.wout = pack_conv_weight(P.raw(p + "conv_out.weight").detach().float()).to(P.dev)
        self.bout = P.f32(p + "conv_out.bias")

    def __call__(self, ops, x, return_fea=False, pool=None):
        h = ops.conv_small_cin(x.float().contiguous(), *self.conv_in)
        fea = []
        for lvl, (blocks, ds) in enumerate(self.down):
            for blk in blocks:
                h = blk(ops, h, pool=pool)
            if return_fea and lvl in (1, 2):
                fea.append(h)
            if ds is not None:
                h = ds(ops, h, pool=pool)
        h = self.mid2(ops, self.attn(ops, self.mid1(ops, h, pool=pool)), pool=pool)
        h = ops.conv3x3_small_cout(_gn_silu(ops, h, self.norm_out, 1e-6, True), self.wout, self.bout)
        return (h, fea) if return_fea else h


def _shapes_resnet(sh, p, ci, co, skip_key="nin_shortcut"):
    sh[p + ".norm1.weight"] = sh[p + ".norm1.bias"] = (ci,)
    sh[p + ".conv1.weight"], sh[p + ".conv1.bias"] = (co, ci, 3, 3), (co,)
    sh[p + ".norm2.weight"] = sh[p + ".norm2.bias"] = (co,)
    sh[p + ".conv2.weight"], sh[p + ".conv2.bias"] = (co, co, 3, 3), (co,)
    if ci != co:
        sh[f"{p}.{skip_key}.weight"], sh[f"{p}.{skip_key}.bias"] = (co, ci, 1, 1), (co,)


def _shapes_attn(sh, p, C):
    sh[p + ".norm.weight"] = sh[p + ".norm.bias"] = (C,)
    for n in ("q", "k", "v", "proj_out"):
        sh[f"{p}.{n}.weight"], sh[f"{p}.{n}.bias"] = (C, C, 1, 1), (C,)


def encoder_shapes(dd, p="encoder."):
    sh, ch, mult, nrb = {}, dd["ch"], dd["ch_mult"], dd["num_res_blocks"]
    sh[p + "conv_in.weight"], sh[p + "conv_in.bias"] = (ch, dd["in_channels"], 3, 3), (ch,)
    cin = ch
    for lvl, m in enumerate(mult):
        for b in range(nrb):
            _shapes_resnet(sh, f"{p}down.{lvl}.block.{b}", cin, ch * m)
            cin = ch * m
        if lvl != len(mult) - 1:
            sh[f"{p}down.{lvl}.downsample.conv.weight"], sh[f"{p}down.{lvl}.downsample.conv.bias"] = (cin, cin, 3, 3), (cin,)
    _shapes_resnet(sh, p + "mid.block_1", cin, cin); _shapes_attn(sh, p + "mid.attn_1", cin)
    _shapes_resnet(sh, p + "mid.block_2", cin, cin)
    sh[p + "norm_out.weight"] = sh[p + "norm_out.bias"] = (cin,)
    zc = 2 * dd["z_channels"] if dd.get("double_z", True) else dd["z_channels"]
    sh[p + "conv_out.weight"], sh[p + "conv_out.bias"] = (zc, cin, 3, 3), (zc,)
    return sh


class AutoencoderKL(_ModuleBase):
    """ldm/models/autoencoder.py:299.  ``encode`` is on the hot path (LR clip -> latent); ``decode`` (:361, reached through
    ``LatentDiffusionVSRTextWT.decode_first_stage``, ddpm.py:3786) is the plain image decoder on the same kernels."""

    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None, ops=None, **ignored):
        self.dd, self.embed_dim = dict(ddconfig), embed_dim
        self.ops = ops or _cuda_ops
        self.pool = StatsPool()
        self.loaded = False

    def expected_shapes(self, with_decoder=True):
        sh = encoder_shapes(self.dd)
        z = self.dd["z_channels"]
        sh["quant_conv.weight"], sh["quant_conv.bias"] = (2 * self.embed_dim, 2 * z, 1, 1), (2 * self.embed_dim,)
        if with_decoder:
            sh.update(decoder_shapes(self.dd, video=False))
            sh["post_quant_conv.weight"], sh["post_quant_conv.bias"] = (z, self.embed_dim, 1, 1), (z,)
        return sh

    def load_state_dict(self, sd, strict=True, device="cuda"):
        """The decoder half is optional (an encoder-only state_dict keeps working; ``decode`` then raises)."""
        P = _Packed(sd, torch.device(device))
        self.encoder = _Encoder(P, "encoder.", self.dd)
        self.qw, self.qb = P.f32("quant_conv.weight"), P.f32("quant_conv.bias")
        has_dec = "decoder.conv_in.weight" in sd
        self.decoder = _VideoDecoder(P, "decoder.", self.dd, video=False) if has_dec else None
        if has_dec:
            self.pqw, self.pqb = P.f32("post_quant_conv.weight"), P.f32("post_quant_conv.bias")
        self.loaded = True
        missing = [k for k in self.expected_shapes(with_decoder=has_dec) if k not in sd]
        unexpected = [k for k in sd if k not in P.used]
        if strict and missing:
            raise KeyError(f"state_dict is missing {missing[:5]}...")
        return missing, unexpected

    def decode(self, z):
        """autoencoder.py:361-364 -> (N,3,H,W) fp32"""
        if getattr(self, "decoder", None) is None:
            raise RuntimeError("AutoencoderKL.decode: the loaded state_dict has no decoder.* / post_quant_conv.* weights")
        self.pool.reset(z.shape[0], z.device)
        self.ops.stats_pool_reset()
        z = self.ops.conv_small_f32(z.float().contiguous(), self.pqw, self.pqb)
        return self.decoder(self.ops, z, None, pool=self.pool)

    def encode(self, x, return_encfea=False):
        """autoencoder.py:347-353"""
        self.pool.reset(x.shape[0], x.device)
        self.ops.stats_pool_reset()
        moments = self.ops.conv_small_f32(self.encoder(self.ops, x, pool=self.pool), self.qw, self.qb)
        post = DiagonalGaussianDistribution(moments, self.ops)
        return (post, moments) if return_encfea else post


class _RDB:
    """ResidualDenseBlock, basicsr/archs/rrdbnet_arch.py:9-39, on a channel-slotted buffer [x | s1 | s2 | s3 | s4]
    (slot = 64 channels, 32 used); conv_k reads x and slots < k, writes slot k.  Weights are re-laid-out to match."""

    def __init__(self, P, p, C, grow=32):
        assert grow == 32
        self.C = C
        self.w, self.b = [], []
        for k in range(1, 6):
            w = P.raw(f"{p}.conv{k}.weight").detach().float()
            co = w.shape[0]
            wp = torch.zeros(co, C + 64 * (k - 1), 3, 3)
            wp[:, :C] = w[:, :C]
            for j in range(k - 1):
                wp[:, C + 64 * j: C + 64 * j + 32] = w[:, C + 32 * j: C + 32 * (j + 1)]
            self.w.append(pack_conv_weight(wp).to(P.dev))
            self.b.append(P.f32(f"{p}.conv{k}.bias"))

    def __call__(self, ops, buf, out):
        """buf: [T,H,W,C+256] with x in [:C] and zeros elsewhere; out: destination view for  0.2*conv5 + x."""
        C = self.C
        for k in range(1, 5):
            ops.conv_gemm(buf[..., :C + 64 * (k - 1)], self.w[k - 1], taps=TAPS_3X3, bias=self.b[k - 1], act=ACT_LRELU02,
                          out=buf, out_col0=C + 64 * (k - 1), block_n=32)
        return ops.conv_gemm(buf, self.w[4], taps=TAPS_3X3, bias=self.b[4], alpha=0.2, beta=1.0, res=buf[..., :C],
                             out=out)


class _FuseBlock:
    """Fuse_sft_block_ResidualDenseBlock, model.py:1354-1367."""

    def __init__(self, P, p, C, num_block=2):
        self.C = C
        self.enc1 = _ResnetBlock(P, p + ".encode_enc_1", skip_key="conv_out")
        self.rdbs = [_RDB(P, f"{p}.encode_enc_2.{i}", C) for i in range(num_block)]
        self.enc3 = _ResnetBlock(P, p + ".encode_enc_3", skip_key="conv_out")

    def __call__(self, ops, enc_feat, dec_feat, w, pool=None):
        T, H, W, C = dec_feat.shape
        bufs = [torch.zeros(T, H, W, C + 256, device=dec_feat.device, dtype=torch.float16) for _ in self.rdbs]
        self.enc1(ops, enc_feat, x2=dec_feat, out=bufs[0])          # writes columns [0, C)
        e = None
        for i, rdb in enumerate(self.rdbs):
            last = i == len(self.rdbs) - 1
            dst = torch.empty(T, H, W, C, device=dec_feat.device, dtype=torch.float16) if last else bufs[i + 1]
            e = rdb(ops, bufs[i], dst)
        e = self.enc3(ops, e, pool=pool)
        return ops.axpby(dec_feat, e, 1.0, float(w))                 # dec + w * enc


class _VideoDecoder:
    """VideoDecoder_Mix, model.py:926-1056; with video=False the plain image Decoder, model.py:575-684 (same topology
    without the SpatialTemporalConv after every ResnetBlock and without the encoder-feature fusion blocks)."""

    def __init__(self, P, p, dd, video=True):
        self.nres, self.nrb = len(dd["ch_mult"]), dd["num_res_blocks"]
        self.fusion_w = 1.0
        self.conv_in = (P.f32(p + "conv_in.weight"), P.f32(p + "conv_in.bias"))
        self.mid1, self.attn, self.mid2 = (_ResnetBlock(P, p + "mid.block_1"), _AttnBlock(P, p + "mid.attn_1"),
                                           _ResnetBlock(P, p + "mid.block_2"))
        ident = lambda ops, h, pool=None: h
        self.tmix = _TemporalConv(P, p + "temporal_mixing", dd["num_frames"]) if video else ident
        self.up = {}
        for lvl in range(self.nres):
            blocks = [(_ResnetBlock(P, f"{p}up.{lvl}.block.{b}"),
                       _TemporalConv(P, f"{p}up.{lvl}.temporal_mixing.{b}", dd["num_frames"]) if video else ident)
                      for b in range(self.nrb + 1)]
            ups = _Upsample(P, f"{p}up.{lvl}.upsample") if lvl != 0 else None
            fuse = None
            if video and lvl != self.nres - 1 and lvl != 0:
                fuse = _FuseBlock(P, f"{p}fusion_layer_{lvl}", dd["ch"] * dd["ch_mult"][lvl])
            self.up[lvl] = (blocks, fuse, ups)
        self.norm_out = P.norm(p + "norm_out")
        self.wout = pack_conv_weight(P.raw(p + "conv_out.weight").detach().float()).to(P.dev)
        self.bout = P.f32(p + "conv_out.bias")

    def __call__(self, ops, z, enc_fea, pool=None):
        h = ops.conv_small_cin(z, *self.conv_in)
        h = self.mid1(ops, h, pool=pool)
        h = self.tmix(ops, h, pool=pool)
        h = self.attn(ops, h)
        h = self.mid2(ops, h, pool=pool)
        for lvl in reversed(range(self.nres)):
            blocks, fuse, ups = self.up[lvl]
            for blk, tm in blocks:
                h = tm(ops, blk(ops, h, pool=pool), pool=pool)
            if fuse is not None:
                h = fuse(ops, as_nhwc_f16(enc_fea[lvl - 1], ops), h, self.fusion_w, pool=pool)
            if ups is not None:
                h = ups(ops, h, pool=pool)
        return ops.conv3x3_small_cout(_gn_silu(ops, h, self.norm_out, 1e-6, True), self.wout, self.bout)


def decoder_shapes(dd, p="decoder.", video=True):
    """state_dict manifest of VideoDecoder_Mix (video=True) or the plain Decoder of AutoencoderKL (video=False)"""
    sh, ch, mult, nrb = {}, dd["ch"], dd["ch_mult"], dd["num_res_blocks"]
    nres = len(mult)
    cin = ch * mult[-1]
    sh[p + "conv_in.weight"], sh[p + "conv_in.bias"] = (cin, dd["z_channels"], 3, 3), (cin,)
    _shapes_resnet(sh, p + "mid.block_1", cin, cin); _shapes_attn(sh, p + "mid.attn_1", cin)
    _shapes_resnet(sh, p + "mid.block_2", cin, cin)

    def tconv(q, C):
        if not video:
            return
        sh[q + ".temporal_conv.weight"], sh[q + ".temporal_conv.bias"], sh[q + ".temporal_alpha"] = (C, C, 3, 1, 1), (C,), (1,)

    tconv(p + "temporal_mixing", cin)
    for lvl in reversed(range(nres)):
        co = ch * mult[lvl]
        if video and lvl != nres - 1 and lvl != 0:
            f = f"{p}fusion_layer_{lvl}"
            _shapes_resnet(sh, f + ".encode_enc_1", 2 * co, co, "conv_out")
            for i in range(2):
                for k in range(1, 6):
                    cout = co if k == 5 else 32
                    sh[f"{f}.encode_enc_2.{i}.conv{k}.weight"] = (cout, co + 32 * (k - 1), 3, 3)
                    sh[f"{f}.encode_enc_2.{i}.conv{k}.bias"] = (cout,)
            _shapes_resnet(sh, f + ".encode_enc_3", co, co, "conv_out")
        for b in range(nrb + 1):
            _shapes_resnet(sh, f"{p}up.{lvl}.block.{b}", cin, co)
            cin = co
            tconv(f"{p}up.{lvl}.temporal_mixing.{b}", co)
        if lvl != 0:
            sh[f"{p}up.{lvl}.upsample.conv.weight"], sh[f"{p}up.{lvl}.upsample.conv.bias"] = (co, co, 3, 3), (co,)
    sh[p + "norm_out.weight"] = sh[p + "norm_out.bias"] = (cin,)
    sh[p + "conv_out.weight"], sh[p + "conv_out.bias"] = (dd["out_ch"], cin, 3, 3), (dd["out_ch"],)
    return sh


class VideoAutoencoderKLResi(_ModuleBase):
    """ldm/models/autoencoder.py:1564 (version 1).  ``lossconfig`` and the other training-only arguments are accepted
    and ignored, so the reference YAML instantiates this class unchanged."""

    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None, fusion_w=1.0, freeze_dec=True, synthesis_data=False,
                 use_usm=False, test_gt=False, version=1, ops=None, **ignored):
        if version != 1:
            raise NotImplementedError("VideoDecoder_MixV2 (version != 1) is not used by the shipped configs")
        self.dd, self.embed_dim = dict(ddconfig), embed_dim
        self.ops = ops or _cuda_ops
        self._fusion_w = fusion_w
        self.pool = StatsPool()
        self.loaded = False

    def expected_shapes(self):
        sh = encoder_shapes(self.dd)
        sh.update(decoder_shapes(self.dd))
        z = self.dd["z_channels"]
        sh["quant_conv.weight"], sh["quant_conv.bias"] = (2 * self.embed_dim, 2 * z, 1, 1), (2 * self.embed_dim,)
        sh["post_quant_conv.weight"], sh["post_quant_conv.bias"] = (z, self.embed_dim, 1, 1), (z,)
        return sh

    def load_state_dict(self, sd, strict=True, device="cuda"):
        P = _Packed(sd, torch.device(device))
        self.encoder = _Encoder(P, "encoder.", self.dd)
        self.decoder = _VideoDecoder(P, "decoder.", self.dd)
        self.decoder.fusion_w = self._fusion_w
        self.qw, self.qb = P.f32("quant_conv.weight"), P.f32("quant_conv.bias")
        self.pqw, self.pqb = P.f32("post_quant_conv.weight"), P.f32("post_quant_conv.bias")
        self.loaded = True
        missing = [k for k in self.expected_shapes() if k not in sd]
        unexpected = [k for k in sd if k not in P.used]   # loss.* etc. (ignored like strict=False does)
        if strict and missing:
            raise KeyError(f"state_dict is missing {missing[:5]}...")
        return missing, unexpected

    def encode(self, x):
        """autoencoder.py:1674-1679 -> (posterior, enc_fea); enc_fea: NCHW-shaped fp16 (channels-last) feature taps."""
        self.pool.reset(x.shape[0], x.device)
        self.ops.stats_pool_reset()
        h, fea = self.encoder(self.ops, x, return_fea=True, pool=self.pool)
        moments = self.ops.conv_small_f32(h, self.qw, self.qb)
        return DiagonalGaussianDistribution(moments, self.ops), [nchw_view(f) for f in fea]

    def decode(self, z, enc_fea):
        """autoencoder.py:1687-1690 -> (T,3,H,W) fp32"""
        self.pool.reset(z.shape[0], z.device)
        self.ops.stats_pool_reset()
        z = self.ops.conv_small_f32(z.float().contiguous(), self.pqw, self.pqb)
        return self.decoder(self.ops, z, enc_fea, pool=self.pool)
