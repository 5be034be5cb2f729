"""GPU dev: where do the attention v3 roles wait?  Per-CTA cycle counters (mgld_attention_set_debug_counters)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgld_vsr_b200 import ops, lib as L
dev = "cuda"
so = L.lib()
so.mgld_attention_set_debug_counters.argtypes = [ctypes.c_void_p]
so.mgld_attention_set_debug_counters.restype = None
names = ["mma.wait_kv", "mma.wait_sfree", "mma.wait_pfull", "mma.loop", "sm0.wait_sfull", "sm0.wait_pv", "sm0.wait_turn", "sm0.exp", "sm0.loop",
         "sm1.wait_sfull", "sm1.wait_pv", "sm1.wait_turn", "sm1.exp", "sm1.loop", "tma.wait_empty", "kernel"]
for (B, N, heads) in [(5, 4096, 5), (5, 1024, 10)]:
    dh = 64; C = heads * dh
    qkv = torch.randn(B * N, 3 * C, device=dev).half()
    kw = dict(batch=B, heads=heads, head_dim=dh, nq=N, nkv=N, scale=dh ** -0.5, q_col0=0, k_col0=C, v_col0=2 * C)
    for _ in range(2): ops.attention(qkv, qkv, qkv, **kw)
    nct = ((N + 255) // 256) * heads * B
    dbg = torch.zeros(nct * 16, dtype=torch.int64, device=dev)
    so.mgld_attention_set_debug_counters(ctypes.c_void_p(dbg.data_ptr()))
    ops.attention(qkv, qkv, qkv, **kw)
    torch.cuda.synchronize()
    so.mgld_attention_set_debug_counters(ctypes.c_void_p(0))
    d = dbg.view(nct, 16).double()
    m = d.mean(0)
    nblk = (N + 127) // 128
    print(f"B{B} N{N} h{heads}: {nct} CTAs, {nblk} key blocks; mean cycles per CTA (per key block in brackets)")
    for n, v in zip(names, m.tolist()):
        print(f"   {n:16s} {v:10.0f}  [{v / nblk:7.0f}]")
