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
for (T, H, W, Ci, Co, taps) in [(5, 8, 8, 1280, 1280, 9), (5, 8, 8, 2560, 1280, 9), (5, 16, 16, 1280, 1280, 9), (5, 16, 16, 2560, 1280, 9),
                               (5, 8, 8, 1280, 1280, 1), (5, 16, 16, 512, 512, 9), (5, 8, 8, 512, 512, 9), (5, 16, 16, 1280, 10240, 1), (5, 8, 8, 1280, 10240, 1)]:
    x = torch.randn(T, H, W, Ci, device=dev).half(); w = (torch.randn(Co, taps * Ci, device=dev) * 0.02).half(); b = torch.randn(Co, device=dev)
    out = torch.empty(T, H, W, Co, device=dev, dtype=torch.float16)
    res = []
    for sp in (False, True):
        ops.SPLIT_K = sp
        n0 = ops.LAUNCHES[0]; ops.conv_gemm(x, w, taps=taps, bias=b, out=out); nl = ops.LAUNCHES[0] - n0
        ms = bench(lambda: ops.conv_gemm(x, w, taps=taps, bias=b, out=out))
        res.append(f"split={sp} launches={nl} {ms*1e3:.1f}us {2.0*T*H*W*Ci*Co*taps/ms/1e9:.0f}TF")
    print(f"T{T} {H}x{W} {Ci}->{Co} taps{taps}: " + " | ".join(res), flush=True)
