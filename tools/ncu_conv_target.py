"""Profiling target: a few heavy conv_gemm launches (UNet level-0 3x3 conv, a 1x1 projection), CTA-pair mode on and off.
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -o gpurun_out/prof_conv python tools/ncu_conv_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgld_vsr_b200 import ops
dev = "cuda"
shapes = [(5, 64, 64, 320, 320, 9), (5, 64, 64, 320, 2560, 1), (5, 32, 32, 1280, 640, 9)]
for pair in ("1", "0"):
    os.environ["MGLD_CONV_PAIR"] = pair
    for (T, H, W, Ci, Co, taps) in shapes:
        x = torch.randn(T, H, W, Ci, device=dev).half(); w = (torch.randn(Co, taps * Ci, device=dev) * 0.02).half()
        b = torch.randn(Co, device=dev)
        out = torch.empty(T, H, W, Co, device=dev, dtype=torch.float16)
        ops.conv_gemm(x, w, taps=taps, bias=b, out=out)
        torch.cuda.synchronize()
print("done")
