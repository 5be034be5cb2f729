"""Clip batching (`(b t)` with b > 1) on CPU through tests/emu_ops.py: several independent clips / UNet tiles / VAE tiles
go through one struct-encoder + UNet evaluation and one lock-step DDPM loop, and every clip must come out exactly as if
it had been processed alone (the script re-seeds per unit, script :428).  The reference's own modules define the `(b t)`
semantics of the temporal layers (util.py:301-310, attention.py:135-141 with oracle patch D1)."""
import torch
import torch.nn.functional as F

import emu_ops
from common import TINY_DD, TINY_STRUCT, TINY_UNET, det_state_dict, det_tensor, rel_err
from oracle import torch_ref as R
from test_host_graph_cpu import _tiny_ldm

T = 2


def test_unet_two_clips_in_one_batch():
    from mgld_vsr_b200.unet import InflatedEncoderUNetModelWT, InflatedUNetModelDualcondV2
    unet = InflatedUNetModelDualcondV2(**TINY_UNET, ops=emu_ops)
    se = InflatedEncoderUNetModelWT(**TINY_STRUCT, ops=emu_ops)
    sd_u, sd_s = det_state_dict(unet.expected_shapes()), det_state_dict(se.expected_shapes())
    unet.load_state_dict(sd_u, device="cpu")
    se.load_state_dict(sd_s, device="cpu")
    x, lat = det_tensor("x2", (2 * T, 4, 32, 32)), det_tensor("lat2", (2 * T, 4, 32, 32))
    ctx, t = det_tensor("ctx", (1, 77, 128)), torch.tensor([321])
    both = unet(x, t, ctx, se(lat, t))
    # (i) the oracle evaluates the same `(b t)` batch with the reference's rearranges (b = 2 clips of num_frames)
    with torch.no_grad():
        ref = R.unet_forward(sd_u, TINY_UNET, x, t, ctx, R.struct_encoder_forward(sd_s, TINY_STRUCT, lat, t, prefix=""),
                             prefix="")
    assert rel_err(both, ref) < 5e-3
    # (ii) clips do not see each other: each half equals the clip run alone (temporal conv padding / temporal attention
    # stay inside a clip)
    for k in range(2):
        alone = unet(x[k * T:(k + 1) * T], t, ctx, se(lat[k * T:(k + 1) * T], t))
        assert rel_err(both[k * T:(k + 1) * T], alone) < 1e-5
    # a clip boundary in the wrong place would change the result: swapping frames across clips must matter
    xs = x.clone()
    xs[[1, 2]] = xs[[2, 1]]
    assert rel_err(unet(xs, t, ctx, se(lat, t))[[0, 3]], both[[0, 3]]) > 1e-4


def _flows(key, h, w):
    ff = 1.5 * F.interpolate(det_tensor(key + "f", (T - 1, 2, 6, 5)), size=(h, w), mode="bicubic")[None]
    fb = -ff + 0.2 * F.interpolate(det_tensor(key + "b", (T - 1, 2, 6, 5)), size=(h, w), mode="bicubic")[None]
    fo, bo = R.forward_backward_consistency_check(fb[:, 0], ff[:, 0])
    return ff, fb, fo[:, None, None], bo[:, None, None]


def test_sample_canvas_num_clips_equals_clip_by_clip():
    m, _ = _tiny_ldm()
    S, h, w = 2, 48, 40                     # 2 x 2 UNet tiles of 32 with overlap 16
    m.respace(S)
    ctx = det_tensor("ctx", (1, 77, 128))
    clips = []
    for k in range(3):
        clips.append((det_tensor(f"lat{k}", (T, 4, h, w)), det_tensor(f"xT{k}", (T, 4, h, w)), _flows(f"c{k}", h, w)))
    kw = dict(cond=ctx, guidance_scale=-10.0, batch_size=T, timesteps=S, time_replace=S, tile_size=32, tile_overlap=16,
              batch_size_sample=1)
    alone = []
    for lat, x_T, (ff, fb, fo, bo) in clips:
        torch.manual_seed(7)                # the script's per-unit re-seed
        alone.append(m.sample_canvas(struct_cond=lat, x_T=x_T, flows=(ff, fb), masks=(fo, bo), **kw))
    for per_call in (1, 2, 4):              # UNet batch smaller than / equal to / larger than the clip count
        m.unet_clips_per_call = per_call
        torch.manual_seed(7)
        got = m.sample_canvas(struct_cond=torch.cat([c[0] for c in clips], 0), x_T=torch.cat([c[1] for c in clips], 0),
                              flows=tuple(torch.cat([c[2][j] for c in clips], 0) for j in (0, 1)),
                              masks=tuple(torch.cat([c[2][j] for c in clips], 0) for j in (2, 3)), num_clips=3, **kw)
        for k in range(3):
            assert rel_err(got[k * T:(k + 1) * T], alone[k]) < 1e-5, (per_call, k)


def _tiny_pipeline(clips_per_batch):
    from mgld_vsr_b200.autoencoder import VideoAutoencoderKLResi
    from mgld_vsr_b200.pipeline import VSRPipeline
    m, _ = _tiny_ldm()
    vq = VideoAutoencoderKLResi(ddconfig=TINY_DD, lossconfig={"target": "torch.nn.Identity"}, embed_dim=4, ops=emu_ops)
    vq.load_state_dict(det_state_dict(vq.expected_shapes()), device="cpu")
    return VSRPipeline(m, vq, ddpm_steps=2, n_frames=T, vqgantile_size=160, vqgantile_stride=96, tile_overlap=8, seed=11,
                       clips_per_batch=clips_per_batch, input_size=128)


def test_pipeline_units_batched_equal_sequential():
    ctx = det_tensor("ctx", (1, 77, 128))
    units = []
    for k, (H, W) in enumerate([(160, 192), (160, 192), (128, 160), (160, 192)]):   # mixed shapes -> grouped by shape
        ff, fb, fo, bo = _flows(f"u{k}", H // 8, W // 8)
        units.append((det_tensor(f"im{k}", (T, 3, H, W)).clamp(-1, 1), ff[0], fb[0], fo[0], bo[0]))
    units.append((units[0][0], None, None, None, None))                              # a unit without guidance
    seq = _tiny_pipeline(1)._sr_units(units, ctx)
    bat = _tiny_pipeline(2)._sr_units(units, ctx)
    for a, b in zip(seq, bat):
        assert a.shape == b.shape and rel_err(a, b) < 1e-5
    assert rel_err(seq[0], seq[1]) > 1e-3                                            # different clips, different output


def test_pipeline_tiled_segment_cut_and_assemble():
    """VAE-tiled branch (script :417-474): the units of a segment, batched, assemble to the sequential result."""
    ctx = det_tensor("ctx", (1, 77, 128))
    H, W = 160, 256                                                                   # > vqgantile_size (160) in width
    seg = det_tensor("seg", (T, 3, H, W)).clamp(-1, 1)
    flows = tuple(f[0] for f in _flows("seg", H // 8, W // 8)[:2])
    outs = []
    for cpb in (1, 2):
        pipe = _tiny_pipeline(cpb)
        pipe.upsample_scale = pipe.upscale
        meta, units = pipe._segment_units(seg, flows_override=flows)
        assert len(units) == 2 and units[0][0].shape == (T, 3, 160, 160) and units[0][1].shape == (T - 1, 2, 20, 20)
        outs.append(pipe._segment_assemble(meta, pipe._sr_units(units, ctx)))
    assert outs[0].shape == (T, 3, H, W) and rel_err(outs[0], outs[1]) < 1e-5


def test_latent_dump_format(tmp_path):
    """`_w_latent` dump (scripts/vsr_val_ddpm_text_T_vqganfin_w_latent.py:396-397): per frame a (4,h,w) float32 .npy of the
    sampled latent, readable with np.load, equal to what sample_canvas returned"""
    import numpy as np
    from mgld_vsr_b200.pipeline import save_latents_npy
    ctx = det_tensor("ctx", (1, 77, 128))
    pipe = _tiny_pipeline(1)
    pipe.keep_latents = True
    unit = (det_tensor("im_lat", (T, 3, 128, 160)).clamp(-1, 1), None, None, None, None)
    pipe._sr_units([unit], ctx)
    lat = torch.cat(pipe.last_latents, 0)
    assert lat.shape == (T, 4, 16, 20)
    paths = save_latents_npy(lat, str(tmp_path / "lat"), [f"{i:08d}" for i in range(T)])
    for i, p in enumerate(paths):
        a = np.load(p)
        assert a.dtype == np.float32 and a.shape == (4, 16, 20) and np.array_equal(a, lat[i].numpy())
