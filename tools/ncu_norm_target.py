"""Profiling target for the normalisation kernels at UNet level-0 / level-2 sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mgld_vsr_b200 import ops
dev = "cuda"
for (T, HW, C) in [(5, 4096, 320), (5, 256, 1280)]:
    x = torch.randn(T, HW, C, device=dev).half(); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    for _ in range(3):
        s = ops.gn_stats(x); y = ops.gn_apply(x, s, 1e-5, g, b, True); z = ops.layernorm(x.reshape(T * HW, C), g, b)
torch.cuda.synchronize(); print("done")
