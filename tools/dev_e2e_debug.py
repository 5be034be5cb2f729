"""development probe (GPU): where does the product pipeline leave the reference-script golden? flows / masks / latents"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from common import *
from oracle import pipeline_ref as PR, torch_ref as R
import test_e2e_gpu as E
from test_reference_pipeline import CASES, T, lr_segment
from mgld_vsr_b200.pipeline import VSRPipeline
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
name = "untiled_adain"
gold = torch.load(os.path.join(GOLDEN, "pipeline.pt"))[name]
S = gold["ddpm_steps"]
Hh, Ww, ts, st, cf, us = CASES[name]
m, vq, sd, vq_sd, ctx, dd = E.build_models(T, S)
pipe = VSRPipeline(m, vq, ddpm_steps=S, n_frames=T, vqgantile_size=ts, vqgantile_stride=st, colorfix_type=cf, seed=42)
pipe.upsample_scale = us
seg = lr_segment(name, Hh, Ww).to("cuda")
im = seg.clamp(-1, 1)
with torch.no_grad():
    oflows, ofo, obo = PR.estimate_flows(sd, im)
pflows, (pfo, pbo) = pipe.estimate_flows(im)
print("flow rel err", rel_err(pflows[0], oflows[0]), rel_err(pflows[1], oflows[1]), "flow absmax", oflows[0].abs().max().item(),
      "mask mismatch", (pfo != ofo).float().mean().item(), (pbo != obo).float().mean().item(), "occ frac", ofo.mean().item())
m.flownet_model.use_cuda_graph = False
pflows2, _ = pipe.estimate_flows(im)
print("no-graph RAFT flow rel err", rel_err(pflows2[0], oflows[0]), rel_err(pflows2[1], oflows[1]))
caps, orig = [], m.sample_canvas
def cap(**kw):
    out = orig(**kw); caps.append((kw, out)); return out
m.sample_canvas = cap
g = gold["units"][0]
for label, fo_ in (("own RAFT flows", None), ("oracle flows", [oflows[0], oflows[1]])):
    for wsg in (True, False):
        m.whole_step_graph = wsg
        caps.clear()
        with cpu_rng():
            sr = pipe.super_resolve_segment(seg, ctx, flows_override=fo_)
        d = (caps[0][1].cpu() - g["samples"]).abs()
        print(label, "whole_step_graph", wsg, "| x_T", rel_err(caps[0][0]["x_T"].cpu(), g["x_T"]), "samples mean", d.mean().item(), "frac>2e-2", (d > 2e-2).float().mean().item(), "max", d.max().item())
# oracle on GPU with the same stream
trace = []
rng = E.DeviceRng(42); rng.seed()
with torch.no_grad():
    ref = PR.sr_segment(sd, TINY_UNET, TINY_STRUCT, dd, vq_sd, dd, seg, ctx, rng, ddpm_steps=S, trace=trace)
d = (trace[0]["samples"].cpu() - g["samples"]).abs()
print("oracle(GPU fp32) vs golden: samples mean", d.mean().item(), "max", d.max().item())
