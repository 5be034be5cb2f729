"""GPU dev perf: conv_gemm TFLOP/s on the UNet's main shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgld_vsr_b200 import ops
dev = "cuda"
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
shapes = [  # T,H,W,Cin,Cout,taps
    (5, 64, 64, 320, 320, 9), (5, 64, 64, 640, 320, 9), (5, 32, 32, 640, 640, 9), (5, 32, 32, 1280, 640, 9),
    (5, 16, 16, 1280, 1280, 9), (5, 16, 16, 2560, 1280, 9), (5, 8, 8, 1280, 1280, 9), (5, 64, 64, 128, 640, 9),
    (5, 64, 64, 320, 960, 1), (5, 64, 64, 320, 2560, 1), (5, 64, 64, 1280, 320, 1), (5, 16, 16, 1280, 10240, 1),
    (5, 512, 512, 128, 128, 9), (5, 256, 256, 256, 256, 9),
]
for (T, H, W, Ci, Co, taps) in shapes:
    x = torch.randn(T, H, W, Ci, device=dev).half(); w = (torch.randn(Co, taps * Ci, device=dev) * 0.02).half()
    b = torch.randn(Co, device=dev)
    out = torch.empty(T, H, W, Co, device=dev, dtype=torch.float16)
    for bn in ([0, 128, 256] if Co % 256 == 0 else [0, 128] if Co % 128 == 0 else [0, 64, 160] if Co % 160 == 0 else [0, 64]):
        ms = bench(lambda: ops.conv_gemm(x, w, taps=taps, bias=b, out=out, block_n=bn))
        fl = 2.0 * T * H * W * Ci * Co * taps
        print(f"T{T} {H}x{W} {Ci}->{Co} taps{taps} bn{bn}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
