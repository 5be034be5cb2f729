"""Profiling target: UNet 64x64 self-attention shape (batch 5 frames x 5 heads, 4096 tokens, head dim 64), 3 launches, then 3 launches of
the cross-attention against the 77 text tokens at the same level."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgld_vsr_b200 import ops
B, N, heads, dh = int(os.environ.get("MGLD_B", "5")), 4096, 5, 64
C = heads * dh
qkv = torch.randn(B * N, 3 * C, device="cuda").half()
for _ in range(3):
    out = ops.attention(qkv, qkv, qkv, batch=B, heads=heads, head_dim=dh, nq=N, nkv=N, scale=dh ** -0.5, q_col0=0, k_col0=C, v_col0=2 * C)
# the 77-key cross-attention of the same level (cross_attention_kv80_kernel): K / V shared by the frames
kv = torch.randn(77, 2 * C, device="cuda").half()
q = qkv[:, :C].contiguous()
for _ in range(3):
    out2 = ops.attention(q, kv, kv, batch=B, heads=heads, head_dim=dh, nq=N, nkv=77, scale=dh ** -0.5, k_col0=0, v_col0=C, kv_batched=False)
torch.cuda.synchronize()
print("done", out.float().abs().max().item(), out2.float().abs().max().item())
