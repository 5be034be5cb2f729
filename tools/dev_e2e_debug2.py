"""development probe (GPU): tiled golden case — per unit, per DDPM step, teacher-forced product vs fp32 oracle; eps per tile"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import numpy as np
from common import *
from oracle import pipeline_ref as PR, torch_ref as R
import test_e2e_gpu as E
from test_reference_pipeline import CASES, T, lr_segment
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
name = "tiled_wavelet_pad"
S = 2
Hh, Ww, ts, st, cf, us = CASES[name]
m, vq, sd, vq_sd, ctx, dd = E.build_models(T, S)
m.whole_step_graph = False
seg = lr_segment(name, Hh, Ww).to("cuda")
trace = []
rng = E.DeviceRng(42)
with torch.no_grad():
    PR.sr_segment(sd, TINY_UNET, TINY_STRUCT, dd, vq_sd, dd, seg, ctx, rng, ddpm_steps=S, vqgantile_size=ts, vqgantile_stride=st, colorfix=cf, trace=trace)
_, resp, use = R.respaced_schedule(ddpm_steps=S)
rm = R.RefModel(sd, dict(TINY_UNET, num_frames=T), dict(TINY_STRUCT, num_frames=T), resp, use, T)
tw = R.gaussian_weights(64, 64, 1).to("cuda")
tile_weights = m._gaussian_weights(64, 64, 1)
for u, tr in enumerate(trace):
    x = tr["x_T"]
    lat = tr["init_latent"]
    h, w = x.shape[-2:]
    offs = R.canvas_tiles(h, w, 64, 32)
    for i in reversed(range(S)):
        t_in = torch.full((1,), m.ori_timesteps[i], device="cuda", dtype=torch.long)
        with torch.no_grad():
            for (ox, oy) in offs:
                eo = rm.eps(x[:, :, oy:oy + 64, ox:ox + 64], t_in, ctx, lat[:, :, oy:oy + 64, ox:ox + 64])
                ep = m._eps(x[:, :, oy:oy + 64, ox:ox + 64].contiguous(), lat[:, :, oy:oy + 64, ox:ox + 64].contiguous(), t_in, ctx)
                d = (ep - eo).abs()
                print(f"unit {u} step {i} tile ({ox},{oy}) eps range {eo.abs().max().item():.2f} max err {d.max().item():.4f} rel {d.max().item() / eo.abs().max().item():.2e} mean {d.mean().item():.2e}")
            xo, _ = rm.p_sample_canvas(x, ctx, lat, i, tr["noises"][i], None, None, -10.0, 64, 32, tw)
        on = m._step_noise; m._step_noise = lambda a, b, n=tr["noises"][i]: n
        ts_ = torch.full((1,), i, device="cuda", dtype=torch.long)
        got = m.p_sample_canvas(x, ctx, lat, ts_, t_replace=t_in, tile_size=64, tile_overlap=32, batch_size=1, tile_weights=tile_weights, _step=i)
        m._step_noise = on
        d = (got - xo).abs()
        idx = d.flatten().argmax().item()
        print(f"unit {u} step {i}: x range {xo.abs().max().item():.1f} max err {d.max().item():.4f} at {np.unravel_index(idx, d.shape) if False else idx} mean {d.mean().item():.2e}")
        x = xo
