"""GPU dev: where does a clip's time go?  CUDA-event timing of the pipeline phases on the bench workload (8-frame 512^2 clip)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench import fast_state_dict, load_cfg, synthetic_clip
from mgld_vsr_b200 import ops
from mgld_vsr_b200.config import instantiate_from_config
from mgld_vsr_b200.pipeline import VSRPipeline
cfg = load_cfg(); dev = torch.device("cuda", 0)
model = instantiate_from_config(cfg.model, device=str(dev)); vq = instantiate_from_config(cfg.video_vae)
sd = {}
for pre, mod, seed in (("model.diffusion_model.", model.model.diffusion_model, 0), ("structcond_stage_model.", model.structcond_stage_model, 1),
                       ("first_stage_model.", model.first_stage_model, 3), ("flownet_model.", model.flownet_model, 4)):
    sd.update({pre + k: v for k, v in fast_state_dict(mod.expected_shapes(), seed).items()})
for k in sd:
    if k.startswith("flownet_model.update_block.flow_head.conv2"): sd[k] = sd[k] * 0.02
model.load_state_dict(sd, strict=False); del sd
vq.load_state_dict(fast_state_dict(vq.expected_shapes(), 2), device=str(dev))
ctx = torch.randn(1, 77, 1024, generator=torch.Generator().manual_seed(1234)).to(dev)
model.cond_stage_model.set_embedding(ctx)
pipe = VSRPipeline(model, vq, ddpm_steps=int(os.environ.get("MGLD_STEPS", "50")), seed=42)
clip = synthetic_clip(42).to(dev)
rec = collections.OrderedDict()
def timed(obj, name, label=None):
    orig = getattr(obj, name)
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = orig(*a, **k); e1.record()
        rec.setdefault(label or name, []).append((e0, e1))
        return out
    setattr(obj, name, f)
timed(pipe, "segments"); timed(pipe, "estimate_flows"); timed(pipe, "_prepare_unit"); timed(pipe, "_finish_unit"); timed(pipe, "_segment_assemble")
timed(model, "sample_canvas"); timed(model, "compute_flow", "  compute_flow (RAFT)"); timed(vq, "encode", "  vq.encode"); timed(vq, "decode", "  vq.decode")
timed(model, "encode_first_stage", "  encode_first_stage")
for it in range(3):
    rec.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = pipe(clip, context=ctx); e1.record()
    torch.cuda.synchronize()
print(f"clip total {e0.elapsed_time(e1):.1f} ms  ({8 / e0.elapsed_time(e1) * 1e3:.2f} frames/s)")
for k, v in rec.items():
    ms = [a.elapsed_time(b) for a, b in v]
    print(f"  {k:28s} n={len(ms):3d}  total {sum(ms):8.2f} ms   each {sum(ms)/len(ms):8.2f} ms")
