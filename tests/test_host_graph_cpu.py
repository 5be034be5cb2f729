"""Host-side logic on CPU: the product's graph (weight packing, layouts, fused-epilogue usage, tile scheduling, schedule
surgery, RNG order) driven through tests/emu_ops.py — a torch evaluation of each C-ABI op — must reproduce the oracle.
No CUDA kernel runs here; the `-m gpu` suite repeats these comparisons with the real kernels."""
import os

import pytest
import torch
import torch.nn.functional as F

import emu_ops
from common import GOLDEN, TINY_DD, TINY_STRUCT, TINY_UNET, det_state_dict, det_tensor, raft_state_dict, rel_err
from oracle import torch_ref as R

T = 2


@pytest.mark.parametrize("fused_stats", [False, True])
def test_unet_and_struct_encoder_host_graph(monkeypatch, fused_stats):
    # fused_stats: the 3x3 convs of the > 16x16 levels hand the GroupNorm sums of their output to the consumer
    # (ops.conv_stats_slot -> conv_gemm stats_out -> gn_apply / gn_finalize) instead of a group_norm call
    monkeypatch.setattr(emu_ops, "FUSED_CONV_STATS", fused_stats)
    from mgld_vsr_b200.unet import InflatedEncoderUNetModelWT, InflatedUNetModelDualcondV2
    unet = InflatedUNetModelDualcondV2(**TINY_UNET, ops=emu_ops)
    se = InflatedEncoderUNetModelWT(**TINY_STRUCT, ops=emu_ops)
    sd_u, sd_s = det_state_dict(unet.expected_shapes()), det_state_dict(se.expected_shapes())
    unet.load_state_dict(sd_u, device="cpu")
    se.load_state_dict(sd_s, device="cpu")
    gold = torch.load(os.path.join(GOLDEN, "tiny_unet.pt"))
    x, lat = det_tensor("x", (T, 4, 32, 32)), det_tensor("lat", (T, 4, 32, 32))
    ctx, t = det_tensor("ctx", (1, 77, 128)), torch.tensor([500])
    sc = {"32": det_tensor("s32", (T, 64, 32, 32)), "16": det_tensor("s16", (T, 64, 16, 16))}
    assert rel_err(unet(x, t, ctx, sc), gold["eps"]) < 5e-3
    feats = se(lat, t)
    for k in gold["struct"]:
        assert feats[k].shape == gold["struct"][k].shape and rel_err(feats[k], gold["struct"][k]) < 5e-3
    assert rel_err(unet(x, t, ctx, feats), gold["eps_chained"]) < 5e-3        # zero-copy NHWC hand-off
    with pytest.raises(KeyError):
        bad = dict(sd_u)
        bad.pop("out.2.weight")
        InflatedUNetModelDualcondV2(**TINY_UNET, ops=emu_ops).load_state_dict(bad, device="cpu")


@pytest.mark.parametrize("fused_stats", [False, True])
def test_vae_host_graph(monkeypatch, fused_stats):
    monkeypatch.setattr(emu_ops, "FUSED_CONV_STATS", fused_stats)
    from mgld_vsr_b200.autoencoder import AutoencoderKL, VideoAutoencoderKLResi
    gold = torch.load(os.path.join(GOLDEN, "tiny_vae.pt"))
    vq = VideoAutoencoderKLResi(ddconfig=TINY_DD, lossconfig={"target": "ldm.modules.losses.LPIPSWithDiscriminator"},
                                embed_dim=4, ops=emu_ops)
    sd = det_state_dict(vq.expected_shapes())
    sd["loss.logvar"] = torch.zeros(())                # training-only keys in real ckpts are tolerated
    missing, unexpected = vq.load_state_dict(sd, device="cpu")
    assert not missing and unexpected == ["loss.logvar"]
    img, z = det_tensor("img", (T, 3, 64, 64)).clamp(-1, 1), det_tensor("z", (T, 4, 8, 8))
    post, fea = vq.encode(img)
    assert rel_err(post.parameters, gold["moments"]) < 5e-3
    assert rel_err(vq.decode(z, fea), gold["dec"]) < 5e-3
    vq.decoder.fusion_w = 0.0                          # --dec_w 0 switches the encoder-feature fusion off
    sd_nofuse = R.video_vae_decode(sd, TINY_DD, z, [f.float() for f in fea], 0.0)
    assert rel_err(vq.decode(z, fea), sd_nofuse) < 5e-3
    kl = AutoencoderKL(ddconfig=TINY_DD, embed_dim=4, ops=emu_ops)
    kl.load_state_dict(det_state_dict(kl.expected_shapes()), device="cpu")
    post = kl.encode(img)
    assert rel_err(post.parameters, gold["kl_moments"]) < 5e-3
    torch.manual_seed(3)
    s1 = post.sample()
    torch.manual_seed(3)
    s2 = R.gaussian_sample(post.parameters, torch.randn(post.mean.shape))     # same CPU RNG draw as distributions.py:36
    assert torch.allclose(s1, s2, atol=1e-6)


def _tiny_ldm():
    from mgld_vsr_b200.config import _wrap
    from mgld_vsr_b200.ddpm import LatentDiffusionVSRTextWT
    cfg = _wrap(dict(
        first_stage_config=dict(target="ldm.models.autoencoder.AutoencoderKL",
                                params=dict(ddconfig=TINY_DD, embed_dim=4, lossconfig=dict(target="torch.nn.Identity"))),
        cond_stage_config=dict(target="ldm.modules.encoders.modules.FrozenOpenCLIPEmbedder", params=dict(freeze=True)),
        structcond_stage_config=dict(target="ldm.modules.diffusionmodules.openaimodel.InflatedEncoderUNetModelWT",
                                     params=TINY_STRUCT),
        unet_config=dict(target="ldm.modules.diffusionmodules.openaimodel.InflatedUNetModelDualcondV2", params=TINY_UNET)))
    m = LatentDiffusionVSRTextWT(**cfg, flownet_config=None, num_frames=T, linear_start=0.00085, linear_end=0.0120,
                                 timesteps=1000, image_size=512, channels=4, scale_factor=0.18215,
                                 conditioning_key="crossattn", time_replace=1000, ops=emu_ops, device="cpu")
    shapes = {}
    for pre, mod in (("model.diffusion_model.", m.model.diffusion_model), ("first_stage_model.", m.first_stage_model),
                     ("structcond_stage_model.", m.structcond_stage_model)):
        shapes.update({pre + k: v for k, v in mod.expected_shapes().items()})
    sd = det_state_dict(shapes)
    missing, _ = m.load_state_dict(sd, strict=False)
    assert missing == []
    return m, sd


def test_sample_canvas_host_logic():
    m, sd = _tiny_ldm()
    S, h, w = 3, 48, 40
    m.respace(S)
    _, resp, use = R.respaced_schedule(ddpm_steps=S)
    ctx, lat, x_T = det_tensor("ctx", (1, 77, 128)), det_tensor("lat", (T, 4, h, w)), det_tensor("xT", (T, 4, h, w))
    ff = 1.5 * F.interpolate(det_tensor("ff", (T - 1, 2, 6, 5)), size=(h, w), mode="bicubic")[None]
    fb = -ff + 0.2 * F.interpolate(det_tensor("fb", (T - 1, 2, 6, 5)), size=(h, w), mode="bicubic")[None]
    fo, bo = R.forward_backward_consistency_check(fb[:, 0], ff[:, 0])
    fo, bo = fo[:, None, None], bo[:, None, None]
    torch.manual_seed(123)
    noises = {i: torch.randn(T, 4, h, w) for i in reversed(range(S))}
    with torch.no_grad():
        ref = R.RefModel(sd, TINY_UNET, TINY_STRUCT, resp, use, T).sample_canvas(
            ctx, lat, x_T, noises, flows=(ff, fb), masks=(fo, bo), guidance_scale=-10.0, tile_size=32, tile_overlap=16)
    torch.manual_seed(123)
    got = m.sample_canvas(cond=ctx, struct_cond=lat, guidance_scale=-10.0, flows=(ff, fb), masks=(fo, bo), batch_size=T,
                          timesteps=S, time_replace=S, x_T=x_T, tile_size=32, tile_overlap=16, batch_size_sample=1)
    assert rel_err(got, ref) < 1e-2
    with pytest.raises(NotImplementedError):
        m.sample_canvas(cond=ctx, struct_cond=lat, batch_size=T, timesteps=S, time_replace=S, x_T=x_T,
                        batch_size_sample=2)


def test_config_plugin_mechanism(tmp_path):
    """the reference YAML (`target:` dotted paths of the reference package) instantiates this package's classes"""
    import yaml
    from mgld_vsr_b200 import config as C
    from mgld_vsr_b200.unet import InflatedUNetModelDualcondV2
    y = {"model": {"target": "ldm.modules.diffusionmodules.openaimodel.InflatedUNetModelDualcondV2", "params": TINY_UNET},
         "loss": {"target": "ldm.modules.losses.LPIPSWithDiscriminator", "params": {"disc_start": 1}},
         "data": {"target": "main.DataModuleFromConfig", "params": {}}}
    p = tmp_path / "c.yaml"
    p.write_text(yaml.safe_dump(y))
    cfg = C.load_config(str(p))
    assert isinstance(C.instantiate_from_config(cfg.model), InflatedUNetModelDualcondV2)
    assert C.instantiate_from_config(cfg.loss) is None and C.instantiate_from_config(cfg["data"]) is None
    assert isinstance(C.instantiate_from_config({"target": "torch.nn.Identity"}), torch.nn.Identity)
    with pytest.raises(KeyError):
        C.instantiate_from_config({"params": {}})
    assert cfg.model.params.model_channels == 64


def test_flow_api_argument_errors():
    from mgld_vsr_b200 import flow
    x, fl = torch.zeros(1, 4, 8, 8), torch.zeros(1, 8, 9, 2)
    with pytest.raises(AssertionError):
        flow.flow_warp(x, fl, ops=emu_ops)                       # arch_util.py:172 shape assert
    with pytest.raises(ValueError):
        flow.resize_flow(torch.zeros(1, 2, 8, 8), "bogus", (4, 4), ops=emu_ops)
    with pytest.raises(AssertionError):
        flow.forward_backward_consistency_check(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8), ops=emu_ops)
    out = flow.flow_warp(torch.ones(1, 2, 8, 8), torch.zeros(1, 8, 8, 2), ops=emu_ops)
    assert torch.allclose(out, torch.ones(1, 2, 8, 8))


def test_raft_host_graph():
    """channel padding (96->128, 324->384), BatchNorm folding, fused z|r GEMM, two-source GRU inputs, last-iteration-only
    convex upsampling: all invisible in the result"""
    from mgld_vsr_b200.raft import RAFT_SR
    gold = torch.load(os.path.join(GOLDEN, "raft.pt"))
    m = RAFT_SR(ops=emu_ops)
    m.load_state_dict(raft_state_dict(m.expected_shapes()), device="cpu")
    a, b = det_tensor("raft_a", (2, 3, 128, 136)).sigmoid(), det_tensor("raft_b", (2, 3, 128, 136)).sigmoid()
    assert rel_err(m(a, b, iters=10), gold["flow"]) < 5e-3
    with pytest.raises(AssertionError):
        m(a, b[:, :, :64])
