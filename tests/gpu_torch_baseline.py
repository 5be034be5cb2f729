"""TEST INFRASTRUCTURE (not collected by pytest): the oracle (oracle/torch_ref.py = the reference's PyTorch ops) timed on the
GPU under fp16 autocast — the reference's own deployment numerics on cuDNN / cuBLAS / SDPA — for one struct-encoder + UNet
tile-step at the SD-2.1 shapes, next to this package's CUDA-graph tile-step.  Context for DESIGN.md, not a bench value.

    python tests/gpu_torch_baseline.py [frames]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench import fast_state_dict, load_cfg
from oracle import torch_ref as R
from mgld_vsr_b200.unet import InflatedEncoderUNetModelWT, InflatedUNetModelDualcondV2

cfg = load_cfg(); dev = "cuda"
mp = cfg.model.params
ucfg, scfg = dict(mp.unet_config.params), dict(mp.structcond_stage_config.params)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 5
sd_u = {k: v.to(dev) for k, v in fast_state_dict(InflatedUNetModelDualcondV2(**ucfg).expected_shapes(), 0).items()}
sd_s = {k: v.to(dev) for k, v in fast_state_dict(InflatedEncoderUNetModelWT(**scfg).expected_shapes(), 1).items()}
x, lat = torch.randn(T, 4, 64, 64, device=dev), torch.randn(T, 4, 64, 64, device=dev)
ctx, t = torch.randn(1, 77, 1024, device=dev), torch.tensor([500], device=dev)

def step():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        feats = R.struct_encoder_forward(sd_s, scfg, lat, t, prefix="")
        return R.unet_forward(sd_u, ucfg, x, t, ctx, feats, prefix="")

for _ in range(3):
    out = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"torch eager (fp16 autocast, cuDNN/cuBLAS/SDPA) struct-enc + UNet tile-step, T={T}: {ms:.2f} ms = {ms / T:.3f} ms/frame "
      f"({4.837 * T / 5 / ms:.3f} PFLOP/s), finite={torch.isfinite(out).all().item()}")
