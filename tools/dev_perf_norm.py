"""GPU dev perf: GroupNorm paths (single-launch vs multi-pass) and LayerNorm at UNet shapes, CUDA-graph timed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgld_vsr_b200 import ops
dev = "cuda"
def bench(fn, n=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (T, HW, C1, C2) in [(5, 4096, 320, 0), (5, 4096, 320, 320), (5, 4096, 640, 320), (5, 1024, 640, 0), (5, 1024, 1280, 640), (5, 256, 1280, 0),
                        (5, 256, 2560, 0), (5, 64, 1280, 0), (5, 64, 2560, 0)]:
    x1 = torch.randn(T, HW, C1, device=dev).half(); x2 = torch.randn(T, HW, C2, device=dev).half() if C2 else None
    g, b = torch.randn(C1 + C2, device=dev), torch.randn(C1 + C2, device=dev)
    t_f = bench(lambda: ops.group_norm(x1, g, b, 1e-5, True, x2=x2))
    def multi():
        s = ops.gn_stats(x1, x2)
        return ops.gn_apply(x1, s, 1e-5, g, b, True, x2=x2)
    t_m = bench(multi)
    mb = T * HW * (C1 + C2) * 2 * 2 / 1e6
    print(f"GN T{T} HW{HW} C{C1}+{C2}: single-launch {t_f:.1f} us ({mb / t_f * 1e-3:.2f} TB/s)   zero+stats+apply {t_m:.1f} us", flush=True)
for (M, C) in [(20480, 320), (5120, 640), (1280, 1280)]:
    x = torch.randn(M, C, device=dev).half(); g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    t = bench(lambda: ops.layernorm(x, g, b))
    print(f"LN M{M} C{C}: {t:.1f} us ({M * C * 4 / 1e6 / t * 1e-3:.2f} TB/s)", flush=True)
