"""TEST INFRASTRUCTURE — oracle restatement of the per-segment body of the reference's inference script
(scripts/vsr_val_ddpm_text_T_vqganfin_oldcanvas_tile.py:375-530) on top of oracle/torch_ref.py.  Functional fp32
PyTorch on whatever device the inputs live on; NOT part of the product (only tests/, smoke() and bench.py's CPU legs
may import oracle/).

Pinning: tests/test_reference_pipeline.py runs the reference's own script text and its own
``LatentDiffusionVSRTextWT.sample_canvas`` / ``VideoAutoencoderKLResi`` (through tests/ref_harness.py) on the same
seeded inputs and compares; tests/golden/pipeline_*.pt hold outputs of that reference run for the GPU box.

Random numbers: the script re-seeds before every unit (:428) and then draws, in this order, the posterior sample of
the LR latent (distributions.py:36), ``randn_like(init_latent)`` (:434) and one ``noise_like`` per DDPM step
(ddpm.py:4404).  ``rng`` reproduces that order: ``rng.seed()`` then ``rng.randn(shape)``.  The untiled branch (:477-518)
does NOT re-seed: its draws continue the stream of the script's single ``seed_everything(opt.seed)`` (:280); callers
that want the first-segment behaviour call ``rng.seed()`` themselves before ``sr_segment``.
"""
import torch
import torch.nn.functional as F

from . import torch_ref as R


class TorchCpuRng:
    """the reference's stream when everything runs on the CPU: one global generator"""

    def __init__(self, seed, device="cpu"):
        self.seed_value, self.device = seed, device

    def seed(self):
        torch.manual_seed(self.seed_value)

    def randn(self, shape):
        return torch.randn(tuple(shape)).to(self.device)


# ---- scripts/util_image.py:686-769 -------------------------------------------------------------------------------
def tile_starts(length, pch, stride):
    """ImageSpliterTh.extract_starts, util_image.py:709-718"""
    if length <= pch:
        return [0]
    out = []
    for s in range(0, length, stride):
        s = length - pch if s + pch > length else s
        if s not in out:
            out.append(s)
    return out


def tile_boxes(height, width, pch, stride):
    """iteration order of ImageSpliterTh.__next__ (:726-750): width index outer, height index inner"""
    hs, ws = tile_starts(height, pch, stride), tile_starts(width, pch, stride)
    return [(h0, h0 + pch, w0, w0 + pch) for w0 in ws for h0 in hs]


def average_tiles(shape, tiles, boxes, like):
    """ImageSpliterTh.update / gather (:752-769): uniform-count averaging"""
    acc, cnt = torch.zeros(shape, dtype=like.dtype, device=like.device), torch.zeros(shape, dtype=like.dtype, device=like.device)
    for t, (h0, h1, w0, w1) in zip(tiles, boxes):
        acc[:, :, h0:h1, w0:w1] += t
        cnt[:, :, h0:h1, w0:w1] += 1
    assert bool((cnt != 0).all())
    return acc / cnt


# ---- scripts/wavelet_color_fix.py ----------------------------------------------------------------------------------
def adain(content, style, eps=1e-5):
    """adaptive_instance_normalization + calc_mean_std, wavelet_color_fix.py:45-71 (unbiased variance + eps)"""
    def ms(f):
        b, c = f.shape[:2]
        flat = f.reshape(b, c, -1)
        return flat.mean(2).reshape(b, c, 1, 1), (flat.var(2) + eps).sqrt().reshape(b, c, 1, 1)
    sm, ss = ms(style)
    cm, cs = ms(content)
    return (content - cm) / cs * ss + sm


def wavelet_fix(content, style, levels=5):
    """wavelet_reconstruction, wavelet_color_fix.py:72-119: high band of the content + coarsest low band of the style"""
    k = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]], dtype=content.dtype, device=content.device) / 16
    k = k[None, None].repeat(3, 1, 1, 1)

    def bands(img):
        high = torch.zeros_like(img)
        for lv in range(levels):
            r = 2 ** lv
            low = F.conv2d(F.pad(img, (r, r, r, r), mode="replicate"), k, groups=3, dilation=r)
            high = high + (img - low)
            img = low
        return high, img
    return bands(content)[0] + bands(style)[1]


# ---- the segment body ------------------------------------------------------------------------------------------------
def estimate_flows(sd, im, flow_fn=None):
    """script :392-416 -> flows [fwd-prop (T-1,2,h,w), bwd-prop], occlusion masks fwd_occs / bwd_occs (T-1,1,h,w)"""
    T, _, H, W = im.shape
    lq01 = torch.clamp((im + 1.0) / 2.0, min=0.0, max=1.0)
    lq01 = F.interpolate(lq01, size=(H // 4, W // 4), mode="bicubic")[None]
    flows = flow_fn(lq01) if flow_fn is not None else R.compute_flow(sd, lq01)
    flows = [R.resize_flow(f[0], H // 8, W // 8) for f in flows]
    fo, bo = [], []
    for i in range(T - 1):
        a, b = R.forward_backward_consistency_check(flows[1][i:i + 1], flows[0][i:i + 1], alpha=0.01, beta=0.5)
        fo.append(a[:, None])
        bo.append(b[:, None])
    return flows, torch.cat(fo, 0), torch.cat(bo, 0)


def sr_unit(ref_model, sd, dd_first, vq_sd, dd_vq, im, flows, masks, ctx, base_sched, rng, ddpm_steps, scale_factor,
            tile_overlap, colorfix, dec_w, trace=None, reseed=True):
    """one VAE tile of one segment, script :428-473.  flows/masks: (1,T-1,2,h,w) / (1,T-1,1,h,w) pairs or None.
    Only the VAE-tiled branch re-seeds per unit (:428); the untiled branch (:477) continues the caller's stream."""
    if reseed:
        rng.seed()
    moments = R.autoencoder_kl_encode(sd, dd_first, im)
    init_latent = scale_factor * R.gaussian_sample(moments, rng.randn(moments[:, :moments.shape[1] // 2].shape))
    noise = rng.randn(init_latent.shape)
    x_T = base_sched["sqrt_alphas_cumprod"][999].to(im.device) * init_latent + \
        base_sched["sqrt_one_minus_alphas_cumprod"][999].to(im.device) * noise               # q_sample_respace, t = 999
    noises = {i: rng.randn(init_latent.shape) for i in reversed(range(ddpm_steps))}
    samples = ref_model.sample_canvas(ctx, init_latent, x_T, noises, flows=flows, masks=masks, guidance_scale=-10.0,
                                      tile_size=64, tile_overlap=tile_overlap)
    if trace is not None:
        trace.append(dict(init_latent=init_latent, x_T=x_T, noises=noises, samples=samples))
    _, fea = R.video_vae_encode(vq_sd, dd_vq, im)
    x = R.video_vae_decode(vq_sd, dd_vq, samples * (1.0 / scale_factor), fea, dec_w)
    if colorfix == "adain":
        x = adain(x, im)
    elif colorfix == "wavelet":
        x = wavelet_fix(x, im)
    return x


def sr_segment(sd, unet_cfg, struct_cfg, dd_first, vq_sd, dd_vq, init_image, ctx, rng, ddpm_steps=50, scale_factor=0.18215,
               vqgantile_size=960, vqgantile_stride=750, tile_overlap=32, colorfix="adain", upscale=4.0,
               upsample_scale=4.0, dec_w=1.0, flow_fn=None, trace=None):
    """script :375-530 for one (T,3,H,W) segment in [-1,1] -> (T,3,H',W') in [0,1]"""
    T = init_image.shape[0]
    base, resp, use = R.respaced_schedule(ddpm_steps=ddpm_steps)
    model = R.RefModel(sd, dict(unet_cfg, num_frames=T), dict(struct_cfg, num_frames=T), resp, use, T)
    im = init_image.clamp(-1.0, 1.0)
    ori_h, ori_w = im.shape[2:]
    flag_pad = not (ori_h % 32 == 0 and ori_w % 32 == 0)
    if flag_pad:                                                                 # :383-387, both dims grow (D13)
        im = F.pad(im, (0, (ori_w // 32 + 1) * 32 - ori_w, 0, (ori_h // 32 + 1) * 32 - ori_h), mode="reflect")
    flows, fwd_occs, bwd_occs = estimate_flows(sd, im, flow_fn)
    args = (model, sd, dd_first, vq_sd, dd_vq)
    kw = dict(ctx=ctx, base_sched=base, rng=rng, ddpm_steps=ddpm_steps, scale_factor=scale_factor,
              tile_overlap=tile_overlap, colorfix=colorfix, dec_w=dec_w, trace=trace)
    if im.shape[2] > vqgantile_size or im.shape[3] > vqgantile_size:            # :418-475
        boxes = tile_boxes(im.shape[2], im.shape[3], vqgantile_size, vqgantile_stride)
        lboxes = tile_boxes(flows[0].shape[2], flows[0].shape[3], vqgantile_size // 8, vqgantile_stride // 8)  # D10
        assert len(boxes) == len(lboxes)
        tiles = []
        for (h0, h1, w0, w1), (a0, a1, b0, b1) in zip(boxes, lboxes):
            cut = lambda t: t[None, :, :, a0:a1, b0:b1]
            tiles.append(sr_unit(*args, im[:, :, h0:h1, w0:w1], (cut(flows[0]), cut(flows[1])),
                                 (cut(fwd_occs), cut(bwd_occs)), **kw))
        x = average_tiles(im.shape, tiles, boxes, im)
    else:                                                                       # :477-518 (D2 resolved)
        x = sr_unit(*args, im, (flows[0][None], flows[1][None]), (fwd_occs[None], bwd_occs[None]), reseed=False, **kw)
    im_sr = torch.clamp((x + 1.0) / 2.0, min=0.0, max=1.0)
    if upsample_scale > upscale:                                                # :520-527
        im_sr = F.interpolate(im_sr, size=(int(im.shape[-2] * upscale / upsample_scale),
                                           int(im.shape[-1] * upscale / upsample_scale)), mode="bicubic")
        im_sr = torch.clamp(im_sr, min=0.0, max=1.0)
    if flag_pad:                                                                # :531-532
        im_sr = im_sr[:, :, :ori_h, :ori_w]
    return im_sr
