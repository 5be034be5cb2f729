"""GPU dev check: attention v3 variants (MGLD_ATTN_EMU = 0/2/4 exponentials per 8 on the FMA pipe) and v2, accuracy against
fp32 SDPA and speed at the UNet / struct-encoder shapes.  Each variant runs in its own process (the choice is cached)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    import torch.nn.functional as F
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
    torch.manual_seed(0)
    for (B, N, heads, qs) in [(2, 300, 3, 1.0), (1, 520, 2, 1.0), (5, 4096, 5, 1.0), (5, 4096, 5, 4.0), (10, 4096, 5, 1.0), (5, 4096, 4, 1.0),
                              (5, 1024, 10, 1.0), (10, 1024, 10, 1.0), (5, 256, 20, 1.0)]:
        dh = 64; C = heads * dh
        qkv = torch.randn(B * N, 3 * C, device=dev)
        qkv[:, :C] *= qs                       # larger logits: exercises the lazy rescale and the polynomial's range
        qkv = qkv.half()
        kw = dict(batch=B, heads=heads, head_dim=dh, nq=N, nkv=N, scale=dh ** -0.5, q_col0=0, k_col0=C, v_col0=2 * C)
        got = ops.attention(qkv, qkv, qkv, **kw)
        x = qkv.float().reshape(B, N, 3, heads, dh)
        q, k, v = [x[:, :, i].transpose(1, 2) for i in range(3)]
        ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, C)
        err = (got.float() - ref).abs().max().item() / ref.abs().max().item()
        rms = ((got.float() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        out = torch.empty_like(got)
        ms = bench(lambda: ops.attention(qkv, qkv, qkv, out=out, **kw))
        fl = 4.0 * B * heads * N * N * dh
        print(f"  self B{B} N{N} h{heads} qscale{qs}: max-rel {err:.2e} rms-rel {rms:.2e}  {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s", flush=True)
    for (B, N, heads) in [(5, 4096, 5), (5, 1024, 10)]:
        dh = 64; C = heads * dh
        q, kv = torch.randn(B * N, C, device=dev).half(), torch.randn(77, 2 * C, device=dev).half()
        kw = dict(batch=B, heads=heads, head_dim=dh, nq=N, nkv=77, scale=dh ** -0.5, k_col0=0, v_col0=C, kv_batched=False)
        got = ops.attention(q, kv, kv, **kw)
        qq = q.float().reshape(B, N, heads, dh).transpose(1, 2)
        kk = kv[:, :C].float().reshape(1, 77, heads, dh).transpose(1, 2).expand(B, -1, -1, -1)
        vv = kv[:, C:].float().reshape(1, 77, heads, dh).transpose(1, 2).expand(B, -1, -1, -1)
        ref = F.scaled_dot_product_attention(qq, kk, vv).transpose(1, 2).reshape(B * N, C)
        err = (got.float() - ref).abs().max().item() / ref.abs().max().item()
        ms = bench(lambda: ops.attention(q, kv, kv, **kw))
        print(f"  cross B{B} N{N} h{heads} nkv77: max-rel {err:.2e}  {ms*1e3:8.1f} us", flush=True)
else:
    for name, env in (("v2", {"MGLD_ATTN_V2": "1"}), ("v3 emu0", {"MGLD_ATTN_EMU": "0"}), ("v3 emu2", {"MGLD_ATTN_EMU": "2"}),
                      ("v3 emu4", {"MGLD_ATTN_EMU": "4"})):
        print(f"== {name}", flush=True)
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=e, timeout=600)
