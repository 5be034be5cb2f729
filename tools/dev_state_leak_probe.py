"""development probe (GPU): the LR latent of the first unit comes out wrong when a tiled clip was processed earlier in the
process and RAFT runs before the encoder.  Toggle suspects by env: MGLD_PDL, PROBE_RAFT_GRAPH, MGLD_WHOLE_STEP, PROBE_SYNC"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from common import *
from oracle import torch_ref as R
import test_e2e_gpu as E
from test_reference_pipeline import CASES, lr_segment
from mgld_vsr_b200.pipeline import VSRPipeline
from mgld_vsr_b200 import ops
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
raft_graph = os.environ.get("PROBE_RAFT_GRAPH", "1") != "0"
# phase 1: the tiled golden case (T = 2)
name = "tiled_wavelet_pad"
Hh, Ww, ts, st, cf, us = CASES[name]
m, vq, sd, vq_sd, ctx, dd = E.build_models(2, 2)
m.flownet_model.use_cuda_graph = raft_graph
pipe = VSRPipeline(m, vq, ddpm_steps=2, n_frames=2, vqgantile_size=ts, vqgantile_stride=st, colorfix_type=cf, seed=42)
pipe.upsample_scale = us
with cpu_rng():
    pipe.super_resolve_segment(lr_segment(name, Hh, Ww).to("cuda"), ctx)
print("pool keys after phase 1:", {k[:2]: (e[1], sum(e[2])) for k, e in ops._sums_pool.bufs.items()})
del m, vq, pipe
# phase 2: the e2e clip (T = 4), RAFT inside
T, S, n = 4, 2, 8
m, vq, sd, vq_sd, ctx, dd = E.build_models(T, S)
m.flownet_model.use_cuda_graph = raft_graph
g = torch.Generator().manual_seed(2024)
hr = F.interpolate(torch.rand(n, 3, 24, 24, generator=g), size=(512, 512), mode="bicubic").clamp(0, 1)
lr = F.interpolate(hr, size=(128, 128), mode="bicubic", antialias=True).clamp(0, 1).to("cuda") * 2 - 1
pipe = VSRPipeline(m, vq, ddpm_steps=S, n_frames=T, seed=42)
segs, _ = pipe.segments(lr)
im = segs[0].clamp(-1, 1)
ref = R.autoencoder_kl_encode(sd, dd, im)
for label in (("after RAFT",) if os.environ.get("PROBE_SKIP_BEFORE") else ("before RAFT", "after RAFT")):
    if label == "after RAFT":
        pipe.estimate_flows(im)
        pipe.estimate_flows(segs[1].clamp(-1, 1))
        if os.environ.get("PROBE_SYNC"):
            torch.cuda.synchronize()
        print("pool keys after RAFT:", {k[:2]: (e[1], sum(e[2])) for k, e in ops._sums_pool.bufs.items()})
    got = m.encode_first_stage(im).parameters
    print(label, "KL encode moments rel err (1st call)", rel_err(got, ref))
    got = m.encode_first_stage(im).parameters
    print(label, "KL encode moments rel err (2nd call)", rel_err(got, ref))
