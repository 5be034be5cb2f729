"""Profiling target: one eager (no CUDA graph) struct-encoder + UNet tile-step at the SD-2.1 shapes (MGLD_T frames, 64x64 latent).
Run under ncu:  ncu ... python tools/ncu_target.py [n_steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench import fast_state_dict, load_cfg
from mgld_vsr_b200.config import instantiate_from_config
cfg = load_cfg(); dev = "cuda"
mp = cfg.model.params
unet = instantiate_from_config(mp.unet_config); se = instantiate_from_config(mp.structcond_stage_config)
unet.load_state_dict(fast_state_dict(unet.expected_shapes(), 0)); se.load_state_dict(fast_state_dict(se.expected_shapes(), 1))
T = int(os.environ.get("MGLD_T", "5"))   # frames in the (b t) batch: 5 = one clip, 10 = the bench's two clips
x = torch.randn(T, 4, 64, 64, device=dev); lat = torch.randn(T, 4, 64, 64, device=dev)
ctx = torch.randn(1, 77, 1024, device=dev); t = torch.tensor([500], device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for k in range(n):
    if k == n - 1:                       # ncu --profile-from-start off: exactly the last tile-step is captured
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    out = unet(x, t, ctx, se(lat, t))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", out.abs().max().item())
