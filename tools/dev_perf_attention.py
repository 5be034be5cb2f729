"""GPU dev perf: attention kernel at the UNet / struct-encoder shapes."""
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
B0 = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for (B, N, heads, dh) in [(B0, 4096, 5, 64), (B0, 4096, 4, 64), (B0, 1024, 10, 64), (B0, 256, 20, 64), (B0, 1024, 4, 64)]:
    C = heads * dh
    qkv = torch.randn(B * N, 3 * C, device=dev).half()
    out = torch.empty(B * N, C, device=dev, dtype=torch.float16)
    ms = bench(lambda: ops.attention(qkv, qkv, qkv, batch=B, heads=heads, head_dim=dh, nq=N, nkv=N, scale=dh ** -0.5, q_col0=0, k_col0=C, v_col0=2 * C, out=out))
    fl = 4.0 * B * heads * N * N * dh
    print(f"self-attn B{B} N{N} h{heads} d{dh}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    kv = torch.randn(77, 2 * C, device=dev).half(); q = qkv[:, :C].contiguous()
    ms = bench(lambda: ops.attention(q, kv, kv, batch=B, heads=heads, head_dim=dh, nq=N, nkv=77, scale=dh ** -0.5, k_col0=0, v_col0=C, kv_batched=False, out=out))
    print(f"cross-attn B{B} N{N} h{heads} d{dh} nkv77: {ms*1e3:.1f} us", flush=True)
