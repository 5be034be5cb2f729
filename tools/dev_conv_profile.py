"""GPU dev: per-layer-shape conv_gemm time inside one eager struct-encoder + UNet tile-step (CUDA events around every call)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench import fast_state_dict, load_cfg
from mgld_vsr_b200.config import instantiate_from_config
from mgld_vsr_b200 import ops
cfg = load_cfg(); dev = "cuda"
mp = cfg.model.params
unet = instantiate_from_config(mp.unet_config); se = instantiate_from_config(mp.structcond_stage_config)
unet.load_state_dict(fast_state_dict(unet.expected_shapes(), 0)); se.load_state_dict(fast_state_dict(se.expected_shapes(), 1))
T = 5
x = torch.randn(T, 4, 64, 64, device=dev); lat = torch.randn(T, 4, 64, 64, device=dev)
ctx = torch.randn(1, 77, 1024, device=dev); t = torch.tensor([500], device=dev)
rec = collections.defaultdict(lambda: [0, 0.0, 0.0])
orig = ops.conv_gemm
def timed(a, w, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = orig(a, w, **kw); e1.record(); torch.cuda.synchronize()
    a2 = kw.get("a2")
    K = w.shape[1]
    shp = tuple(a.shape[:-1])
    M = 1
    for d in shp: M *= d
    key = (shp, K, w.shape[0], kw.get("taps", 1), kw.get("epilogue", 0), kw.get("act", 0), kw.get("res") is not None, a2 is not None)
    r = rec[key]; r[0] += 1; r[1] += e0.elapsed_time(e1) * 1e3; r[2] += 2.0 * M * K * w.shape[0]
    return out
for it in range(2):
    if it == 1: ops.conv_gemm = timed
    import mgld_vsr_b200.unet as U
    out = unet(x, t, ctx, se(lat, t))
torch.cuda.synchronize()
tot = sum(r[1] for r in rec.values())
print(f"total conv_gemm (incl. split finalize) {tot/1e3:.2f} ms over {sum(r[0] for r in rec.values())} calls")
for k, r in sorted(rec.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{str(k):80s} n={r[0]:3d} {r[1]:8.1f} us  {r[1]/r[0]:7.1f} us/call  {r[2]/r[1]/1e6:7.1f} TFLOP/s  {100*r[1]/tot:4.1f}%")
