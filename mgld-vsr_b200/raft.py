"""RAFT optical flow with the reference's class name, constructor and ``forward(ref, sup, iters=10)`` signature
(``RAFT_SR``, basicsr/archs/raft_arch.py:668-808, model='normal'), on the mgld kernels.

Mapping of the network onto the kernels:
  * 7x7 stems (3 -> 64 stride 2; 2 -> 128 on the flow) ............ mgld_conv_direct_f32
  * 3x3 / 1x1 / (1,5) / (5,1) convolutions ........................ mgld_conv_gemm (tcgen05 implicit GEMM); stride-2 3x3 via
                                                                    mgld_im2col_s2, stride-2 1x1 via mgld_subsample2
  * InstanceNorm (fnet) ........................................... mgld_gn_stats_f16(groups=C) + mgld_instance_norm_apply_f16
  * BatchNorm (cnet, eval) ........................................ folded into the conv weights at load time
  * all-pairs correlation fmap1^T fmap2 / sqrt(256) ................ mgld_conv_gemm per pair (fp32 out), pyramid by
                                                                    mgld_avgpool2_f32, 4x9x9 window by mgld_corr_lookup_f32
  * SepConvGRU: z|r in one GEMM with a sigmoid epilogue, q with tanh, gating by mgld_gru_rh / mgld_gru_update;
    torch.cat([h, x]) is the two-source A operand of the GEMM (no concat is materialised)
  * convex 8x upsampling ........................................... mgld_convex_upsample8_f32 (only after the LAST iteration:
                                                                    the reference recomputes and discards it every iteration)
96-channel layers are zero-padded to 128 channels and the 324 correlation channels to 384 (K must be a multiple of 64);
the padding channels are exact zeros end to end, so results are unchanged.
"""
import torch
import torch.nn.functional as F

from . import ops as _cuda_ops
from .ops import (ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, TAPS_1, TAPS_1X5, TAPS_3X3, TAPS_5X1, pack_conv_weight)
from .unet import _ModuleBase


def _pad_w(w, ci_pad=None, co_pad=None):
    co, ci = w.shape[:2]
    ci_pad, co_pad = ci_pad or ci, co_pad or co
    out = torch.zeros(co_pad, ci_pad, *w.shape[2:], dtype=w.dtype)
    out[:co, :ci] = w
    return out


def _pad_b(b, co_pad):
    out = torch.zeros(co_pad, dtype=b.dtype)
    out[:b.shape[0]] = b
    return out


def _c(ch):
    """channel count as laid out in memory (96 -> 128)"""
    return 128 if ch == 96 else ch


class _Conv:
    def __init__(self, sd, p, dev, bn=None, ci_pad=None, co_pad=None, taps=None):
        w, b = sd[p + ".weight"].detach().float(), sd[p + ".bias"].detach().float()
        if bn is not None:   # fold eval-mode BatchNorm2d (raft_arch.py:104-108): y = (conv - mean) * g / sqrt(var + eps) + beta
            g, beta = sd[bn + ".weight"].float(), sd[bn + ".bias"].float()
            mean, var = sd[bn + ".running_mean"].float(), sd[bn + ".running_var"].float()
            s = g / torch.sqrt(var + 1e-5)
            w, b = w * s[:, None, None, None], (b - mean) * s + beta
        kh, kw = w.shape[2:]
        self.taps = taps or {(1, 1): TAPS_1, (3, 3): TAPS_3X3, (1, 5): TAPS_1X5, (5, 1): TAPS_5X1}[(kh, kw)]
        self.w = pack_conv_weight(_pad_w(w, ci_pad, co_pad)).to(dev)
        self.b = _pad_b(b, co_pad or w.shape[0]).to(dev)

    def __call__(self, ops, x, **kw):
        return ops.conv_gemm(x, self.w, taps=self.taps, bias=self.b, **kw)


class _ResidualBlock:
    """ResidualBlock, raft_arch.py:95-136"""

    def __init__(self, sd, p, dev, cin, cout, kind, stride):
        self.kind, self.stride = kind, stride
        bn = (lambda n: f"{p}.{n}") if kind == "batch" else (lambda n: None)
        self.c1 = _Conv(sd, p + ".conv1", dev, bn("norm1"), _c(cin), _c(cout), taps=TAPS_3X3 if stride == 1 else TAPS_1)
        self.c2 = _Conv(sd, p + ".conv2", dev, bn("norm2"), _c(cout), _c(cout))
        # the reference registers norm3 both as `norm3` and as `downsample.1`
        self.ds = _Conv(sd, p + ".downsample.0", dev, bn("downsample.1"), _c(cin), _c(cout)) if stride != 1 else None

    def _norm_relu(self, ops, y):
        return ops.instance_norm(y, relu=True) if self.kind == "instance" else y

    def __call__(self, ops, x):
        act = ACT_NONE if self.kind == "instance" else ACT_RELU
        src = x if self.stride == 1 else ops.im2col_s2(x, 1)
        y = self._norm_relu(ops, self.c1(ops, src, act=act))
        y = self._norm_relu(ops, self.c2(ops, y, act=act))
        if self.ds is not None:
            x = self.ds(ops, ops.subsample2(x))
            if self.kind == "instance":
                x = ops.instance_norm(x, relu=False)
        return ops.axpby(x, y, 1.0, 1.0, relu=True)


class _Encoder:
    """BasicEncoder, raft_arch.py:199-268"""

    def __init__(self, sd, p, dev, kind, out_dim):
        self.kind = kind
        w, b = sd[p + ".conv1.weight"].detach().float(), sd[p + ".conv1.bias"].detach().float()
        if kind == "batch":
            g, beta = sd[p + ".norm1.weight"].float(), sd[p + ".norm1.bias"].float()
            s = g / torch.sqrt(sd[p + ".norm1.running_var"].float() + 1e-5)
            w, b = w * s[:, None, None, None], (b - sd[p + ".norm1.running_mean"].float()) * s + beta
        self.stem = (w.contiguous().to(dev), b.to(dev))
        self.blocks, cin = [], 64
        for li, (dim, stride) in enumerate(((64, 1), (96, 2), (128, 2)), start=1):
            self.blocks.append(_ResidualBlock(sd, f"{p}.layer{li}.0", dev, cin, dim, kind, stride))
            self.blocks.append(_ResidualBlock(sd, f"{p}.layer{li}.1", dev, dim, dim, kind, 1))
            cin = dim
        self.out_w = sd[p + ".conv2.weight"].detach().float()
        self.out_b = sd[p + ".conv2.bias"].detach().float()
        self.dev = dev

    def trunk(self, ops, x):
        if self.kind == "instance":
            h = ops.instance_norm(ops.conv_direct(x, *self.stem, stride=2, pad=3, relu=False), relu=True)
        else:
            h = ops.conv_direct(x, *self.stem, stride=2, pad=3, relu=True)
        for blk in self.blocks:
            h = blk(ops, h)
        return h


class RAFT_SR(_ModuleBase):
    def __init__(self, model="normal", load_path=None, ops=None, **ignored):
        if model != "normal":
            raise NotImplementedError("only the 'normal' RAFT of the shipped config is implemented")
        self.ops = ops or _cuda_ops
        self.hidden_dim = self.context_dim = 128
        self.corr_levels, self.corr_radius = 4, 4
        self.load_path = load_path          # the reference loads raft-things.pth here; use load_state_dict instead
        self.loaded = False

    # ---- state_dict manifest (raft_arch.py:684-696) ------------------------------------------------------------------
    def expected_shapes(self):
        sh = {}

        def conv(p, co, ci, kh, kw=None):
            sh[p + ".weight"], sh[p + ".bias"] = (co, ci, kh, kw or kh), (co,)

        def bnorm(p, c):
            for k in ("weight", "bias", "running_mean", "running_var"):
                sh[f"{p}.{k}"] = (c,)
            sh[p + ".num_batches_tracked"] = ()

        for enc, kind, od in (("fnet", "instance", 256), ("cnet", "batch", 256)):
            conv(enc + ".conv1", 64, 3, 7)
            if kind == "batch":
                bnorm(enc + ".norm1", 64)
            cin = 64
            for li, (dim, stride) in enumerate(((64, 1), (96, 2), (128, 2)), start=1):
                for bi, (ci, st) in enumerate(((cin, stride), (dim, 1))):
                    p = f"{enc}.layer{li}.{bi}"
                    conv(p + ".conv1", dim, ci, 3); conv(p + ".conv2", dim, dim, 3)
                    if kind == "batch":
                        bnorm(p + ".norm1", dim); bnorm(p + ".norm2", dim)
                    if st != 1:
                        conv(p + ".downsample.0", dim, ci, 1)
                        if kind == "batch":
                            bnorm(p + ".norm3", dim); bnorm(p + ".downsample.1", dim)
                cin = dim
            conv(enc + ".conv2", od, 128, 1)
        u = "update_block"
        conv(u + ".encoder.convc1", 256, 324, 1); conv(u + ".encoder.convc2", 192, 256, 3)
        conv(u + ".encoder.convf1", 128, 2, 7); conv(u + ".encoder.convf2", 64, 128, 3)
        conv(u + ".encoder.conv", 126, 256, 3)
        for g in "zrq":
            conv(f"{u}.gru.conv{g}1", 128, 384, 1, 5); conv(f"{u}.gru.conv{g}2", 128, 384, 5, 1)
        conv(u + ".flow_head.conv1", 256, 128, 3); conv(u + ".flow_head.conv2", 2, 256, 3)
        conv(u + ".mask.0", 256, 128, 3); conv(u + ".mask.2", 576, 256, 1)
        return sh

    def load_state_dict(self, sd, strict=True, device="cuda"):
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}     # raft_arch.py:700-703
        dev = torch.device(device)
        self.__dict__.pop("_graphed", None)        # a captured graph replays against the old weight tensors
        missing = [k for k in self.expected_shapes() if k not in sd]
        if strict and missing:
            raise KeyError(f"state_dict is missing {missing[:5]}...")
        self.fnet = _Encoder(sd, "fnet", dev, "instance", 256)
        self.cnet = _Encoder(sd, "cnet", dev, "batch", 256)
        self.f_out = (pack_conv_weight(self.fnet.out_w).to(dev), self.fnet.out_b.to(dev))
        cw, cb = self.cnet.out_w, self.cnet.out_b
        self.c_net = (pack_conv_weight(cw[:128]).to(dev), cb[:128].to(dev))               # -> tanh  (hidden state)
        self.c_inp = (pack_conv_weight(cw[128:]).to(dev), cb[128:].to(dev))               # -> relu  (context)
        u = "update_block"
        self.convc1 = _Conv(sd, u + ".encoder.convc1", dev, ci_pad=384)
        self.convc2 = _Conv(sd, u + ".encoder.convc2", dev)
        self.convf1 = (sd[u + ".encoder.convf1.weight"].detach().float().contiguous().to(dev),
                       sd[u + ".encoder.convf1.bias"].detach().float().to(dev))
        self.convf2 = _Conv(sd, u + ".encoder.convf2", dev)
        self.convm = _Conv(sd, u + ".encoder.conv", dev, co_pad=128)
        self.gru = []
        for sfx in ("1", "2"):
            wz, wr = sd[f"{u}.gru.convz{sfx}.weight"].detach().float(), sd[f"{u}.gru.convr{sfx}.weight"].detach().float()
            bz, br = sd[f"{u}.gru.convz{sfx}.bias"].detach().float(), sd[f"{u}.gru.convr{sfx}.bias"].detach().float()
            taps = TAPS_1X5 if sfx == "1" else TAPS_5X1
            wzr = pack_conv_weight(torch.cat([wz, wr], 0)).to(dev)
            self.gru.append((taps, wzr, torch.cat([bz, br]).to(dev), _Conv(sd, f"{u}.gru.convq{sfx}", dev)))
        self.fh1 = _Conv(sd, u + ".flow_head.conv1", dev)
        self.fh2 = (pack_conv_weight(sd[u + ".flow_head.conv2.weight"].detach().float()).to(dev),
                    sd[u + ".flow_head.conv2.bias"].detach().float().to(dev))
        self.mask0 = _Conv(sd, u + ".mask.0", dev)
        self.mask2 = _Conv(sd, u + ".mask.2", dev)
        self.loaded = True
        return missing, []

    # ---- forward (raft_arch.py:733-808) ---------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, ref, sup, iters=10, flow_init=None, upsample=True):
        """The ~700 launches of one call are tiny (h/8 x w/8 maps): on CUDA they are replayed from a graph captured per
        input shape (`use_cuda_graph`, default on)."""
        assert self.loaded, "load_state_dict() first"
        assert ref.size() == sup.size()                                   # raft_arch.py:800
        assert flow_init is None
        if ref.is_cuda and getattr(self, "use_cuda_graph", True) and hasattr(self.ops, "GraphedFn"):
            runner = self.__dict__.get("_graphed")
            if runner is None or runner[0] != iters:
                fn = self.ops.GraphedFn(lambda a, b: self._forward(a, b, iters))
                runner = self.__dict__["_graphed"] = (iters, fn)
            return runner[1](ref.float().contiguous(), sup.float().contiguous())
        return self._forward(ref, sup, iters)

    def _forward(self, ref, sup, iters=10):
        ops = self.ops
        ops.stats_pool_reset()
        ht, wd = ref.shape[-2:]
        pad_ht, pad_wd = (((ht // 8) + 1) * 8 - ht) % 8, (((wd // 8) + 1) * 8 - wd) % 8
        pad = [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]       # InputPadder 'sintel'
        im1, im2 = F.pad(ref.float(), pad, mode="replicate"), F.pad(sup.float(), pad, mode="replicate")
        B, _, H, W = im1.shape
        h, w = H // 8, W // 8
        M = h * w
        # feature network on both images, context network on the first
        f = self.fnet.trunk(ops, torch.cat([im1, im2], 0).contiguous())
        fmap = ops.conv_gemm(f, self.f_out[0], bias=self.f_out[1])                      # [2B, h, w, 256]
        c = self.cnet.trunk(ops, im1.contiguous())
        net = ops.conv_gemm(c, self.c_net[0], bias=self.c_net[1], act=ACT_TANH)         # [B, h, w, 128]
        xbuf = torch.zeros(B, h, w, 256, device=ref.device, dtype=torch.float16)       # [inp | motion(126) flow(2)]
        ops.conv_gemm(c, self.c_inp[0], bias=self.c_inp[1], act=ACT_RELU, out=xbuf, out_col0=0)
        # all-pairs correlation (fp32) + pyramid
        f1, f2 = fmap[:B].reshape(B, M, 256), fmap[B:].reshape(B, M, 256)
        ldc = (M + 3) // 4 * 4                                                          # 16-byte aligned fp32 rows for the TMA store
        corr = torch.empty(B * M, ldc, device=ref.device, dtype=torch.float32)
        for b in range(B):
            ops.conv_gemm(f1[b], f2[b], alpha=1.0 / 16.0, out_f32=True, out=corr[b * M:(b + 1) * M])
        corr = (corr if ldc == M else corr[:, :M].contiguous()).reshape(B * M, h, w)
        levels = [corr]
        for _ in range(self.corr_levels - 1):
            levels.append(ops.avgpool2_f32(levels[-1]))
        ys, xs = torch.meshgrid(torch.arange(h, device=ref.device), torch.arange(w, device=ref.device), indexing="ij")
        coords0 = torch.stack([xs, ys], dim=0).float()[None].repeat(B, 1, 1, 1)
        flow = torch.zeros(B, 2, h, w, device=ref.device, dtype=torch.float32)
        cfeat = torch.zeros(B, h, w, 384, device=ref.device, dtype=torch.float16)       # 324 used, rest stays zero
        for _ in range(iters):
            ops.corr_lookup(levels, coords0 + flow, cfeat)
            cor = self.convc2(ops, self.convc1(ops, cfeat, act=ACT_RELU), act=ACT_RELU)               # [B,h,w,192]
            flo = self.convf2(ops, ops.conv_direct(flow, *self.convf1, stride=1, pad=3, relu=True), act=ACT_RELU)
            self.convm(ops, cor, a2=flo, act=ACT_RELU, out=xbuf, out_col0=128)                        # cols 128..255
            ops.set_channels(flow, xbuf, 254)                                                           # cat([out, flow])
            for taps, wzr, bzr, convq in self.gru:
                zr = ops.conv_gemm(net, wzr, taps=taps, a2=xbuf, bias=bzr, act=ACT_SIGMOID)            # [B,h,w,256]
                q = convq(ops, ops.gru_rh(zr, net), a2=xbuf, act=ACT_TANH)
                ops.gru_update(zr, q, net)
            delta = ops.conv3x3_small_cout(self.fh1(ops, net, act=ACT_RELU), self.fh2[0], self.fh2[1])  # (B,2,h,w)
            flow = flow + delta
        mask = self.mask2(ops, self.mask0(ops, net, act=ACT_RELU), alpha=0.25)
        flow_up = ops.convex_upsample8(mask, flow)
        h2, w2 = flow_up.shape[-2:]
        return flow_up[..., pad[2]:h2 - pad[3], pad[0]:w2 - pad[1]]

    __call__ = forward
