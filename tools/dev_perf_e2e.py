"""GPU dev perf: full-size (SD-2.1 shape) struct encoder + UNet tile-step, VAE encode/decode at 512^2, kernel breakdown."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import det_state_dict, det_tensor
from mgld_vsr_b200.unet import InflatedUNetModelDualcondV2, InflatedEncoderUNetModelWT
from mgld_vsr_b200.autoencoder import VideoAutoencoderKLResi, AutoencoderKL
dev = "cuda"
UNET = dict(num_frames=5, image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
            num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True,
            use_linear_in_transformer=True, transformer_depth=1, context_dim=1024, use_checkpoint=False, legacy=False, semb_channels=256)
STRUCT = dict(num_frames=5, image_size=96, in_channels=4, model_channels=256, out_channels=256, num_res_blocks=2,
              attention_resolutions=[4, 2, 1], dropout=0, channel_mult=[1, 1, 2, 2], conv_resample=True, dims=2, use_checkpoint=False,
              use_fp16=False, num_heads=4, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
              resblock_updown=False, use_new_attention_order=False)
DD = dict(double_z=True, num_frames=5, z_channels=4, resolution=512, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
          num_res_blocks=2, attn_resolutions=[], dropout=0.0)

def fast_sd(shapes):
    g = torch.Generator().manual_seed(0); sd = {}
    for k, s in shapes.items():
        s = tuple(s)
        if k.endswith("temporal_alpha"): sd[k] = torch.full(s, 0.5)
        elif len(s) == 1 and k.endswith(".weight"): sd[k] = torch.ones(s)
        elif k.endswith(".bias"): sd[k] = torch.zeros(s)
        else:
            fan = 1
            for d in s[1:]: fan *= d
            sd[k] = torch.randn(s, generator=g) * fan ** -0.5
    return sd

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

what = sys.argv[1:] or ["unet", "vae"]
T = int(os.environ.get("MGLD_T", "5"))   # frames per call (5 = one segment; 10 = two segments batched)
if "unet" in what:
    t0 = time.time()
    unet = InflatedUNetModelDualcondV2(**UNET); se = InflatedEncoderUNetModelWT(**STRUCT)
    unet.load_state_dict(fast_sd(unet.expected_shapes())); se.load_state_dict(fast_sd(se.expected_shapes()))
    print(f"weights built+packed in {time.time()-t0:.1f}s", flush=True)
    x = torch.randn(T, 4, 64, 64, device=dev); lat = torch.randn(T, 4, 64, 64, device=dev)
    ctx = torch.randn(1, 77, 1024, device=dev); t = torch.tensor([500], device=dev)
    feats = se(lat, t)
    ms_se = timeit(lambda: se(lat, t))
    ms_un = timeit(lambda: unet(x, t, ctx, feats))
    print(f"eager: struct-enc {ms_se:.2f} ms, unet {ms_un:.2f} ms  ({4.837*T/5:.3f} TFLOP -> {4.837*T/5/(ms_se+ms_un):.2f} PFLOP/s eff)", flush=True)
    # CUDA graph
    def step(): return unet(x, t, ctx, se(lat, t))
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    ms_g = timeit(lambda: g.replay(), n=10)
    print(f"graph: struct-enc+unet tile-step T={T}: {ms_g:.2f} ms -> {4.837*T/5/ms_g:.3f} PFLOP/s, {ms_g/T:.3f} ms/frame", flush=True)
    print("eps finite:", torch.isfinite(out).all().item(), out.abs().max().item())
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    del unet, se, g
    torch.cuda.empty_cache()
if "vae" in what:
    vq = VideoAutoencoderKLResi(ddconfig=DD, lossconfig={"target": "torch.nn.Identity"}, embed_dim=4)
    vq.load_state_dict(fast_sd(vq.expected_shapes()))
    img = torch.rand(T, 3, 512, 512, device=dev) * 2 - 1; z = torch.randn(T, 4, 64, 64, device=dev)
    post, fea = vq.encode(img)
    ms_e = timeit(lambda: vq.encode(img), n=3, warm=1)
    dec = vq.decode(z, fea)
    ms_d = timeit(lambda: vq.decode(z, fea), n=3, warm=1)
    print(f"VAE eager: encode {ms_e:.1f} ms (5.58 TFLOP -> {5.583/ms_e:.3f} PF/s), decode {ms_d:.1f} ms (20.7 TFLOP -> {20.695/ms_d:.3f} PF/s)", flush=True)
    print("dec finite:", torch.isfinite(dec).all().item(), dec.abs().max().item())
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        vq.decode(z, fea); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=15, max_name_column_width=60))
