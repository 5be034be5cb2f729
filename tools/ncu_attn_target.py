"""Profiling target: UNet 64x64 self-attention shape (batch 5 frames x 5 heads, 4096 tokens, head dim 64), 3 launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgld_vsr_b200 import ops
B, N, heads, dh = int(os.environ.get("MGLD_B", "5")), 4096, 5, 64
C = heads * dh
qkv = torch.randn(B * N, 3 * C, device="cuda").half()
for _ in range(3):
    out = ops.attention(qkv, qkv, qkv, batch=B, heads=heads, head_dim=dh, nq=N, nkv=N, scale=dh ** -0.5, q_col0=0, k_col0=C, v_col0=2 * C)
torch.cuda.synchronize()
print("done", out.float().abs().max().item())
