"""GPU dev check: mgld_conv_gemm against torch fp32 ops on fp16-rounded inputs.  Usage: python tools/dev_check_conv_gemm.py [case ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mgld_vsr_b200 import ops

torch.manual_seed(0)
dev = "cuda"

def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).half()

def report(name, got, ref):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-12
    ok = err / den < 4e-3
    print(f"{'PASS' if ok else 'FAIL'} {name}: max_abs_err={err:.4e} ref_max={den:.3e} rel={err/den:.3e}", flush=True)
    return ok

def case_gemm(M, K, N, block_n=0, bias=True, act=ops.ACT_NONE, res=False, f32=False):
    a = rnd(M, K); w = rnd(N, K, scale=K ** -0.5)
    b = torch.randn(N, device=dev) if bias else None
    r = rnd(M, N) if res else None
    out = ops.conv_gemm(a, w, bias=b, act=act, res=r, alpha=0.5 if res else 1.0, beta=2.0 if res else 0.0, block_n=block_n, out_f32=f32)
    ref = a.float() @ w.float().t()
    if bias: ref = ref + b
    if act == ops.ACT_SILU: ref = F.silu(ref)
    if act == ops.ACT_RELU: ref = F.relu(ref)
    if act == ops.ACT_GELU: ref = F.gelu(ref)
    if res: ref = 0.5 * ref + 2.0 * r.float()
    return report(f"gemm M{M} K{K} N{N} bn{block_n} act{act} res{res} f32{f32}", out, ref)

def case_conv3(T, H, W, Cin, Cout, two=False, block_n=0):
    x = rnd(T, H, W, Cin)
    wt = rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5)
    b = torch.randn(Cout, device=dev)
    if two:
        c1 = (Cin // 128) * 64
        a, a2 = x[..., :c1].contiguous(), x[..., c1:].contiguous()
    else:
        a, a2 = x, None
    out = ops.conv_gemm(a, ops.pack_conv_weight(wt), taps=9, a2=a2, bias=b, block_n=block_n)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=1).permute(0, 2, 3, 1)
    return report(f"conv3x3 T{T} {H}x{W} {Cin}->{Cout} two{two} bn{block_n}", out, ref)

def case_t3(T, H, W, C):
    x = rnd(T, H, W, C)
    wt = rnd(C, C, 3, 1, 1, scale=(3 * C) ** -0.5)
    b = torch.randn(C, device=dev)
    alpha = 0.3
    out = ops.conv_gemm(x, ops.pack_temporal_weight(wt), taps=3, bias=b, alpha=alpha, beta=1 - alpha, res=x)
    x5 = x.float().permute(3, 0, 1, 2)[None]  # 1 C T H W
    ref = F.conv3d(x5, wt.float(), b, padding=(1, 0, 0))
    ref = alpha * ref + (1 - alpha) * x5
    ref = ref[0].permute(1, 2, 3, 0)
    return report(f"temporal3 T{T} {H}x{W} C{C}", out, ref)

def case_geglu(M, K, Ni):
    a = rnd(M, K); w = rnd(2 * Ni, K, scale=K ** -0.5); b = torch.randn(2 * Ni, device=dev)
    wp = ops.interleave_pair(w[:Ni], w[Ni:]); bp = ops.interleave_pair(b[:Ni], b[Ni:])
    out = ops.conv_gemm(a, wp, bias=bp, epilogue=ops.EPI_GEGLU)
    y = a.float() @ w.float().t() + b
    ref = y[:, :Ni] * F.gelu(y[:, Ni:])
    return report(f"geglu M{M} K{K} Ni{Ni}", out, ref)

def case_spade(T, H, W, Ch, C):
    actv = rnd(T, H, W, Ch)
    wg = rnd(C, Ch, 3, 3, scale=(9 * Ch) ** -0.5); wb = rnd(C, Ch, 3, 3, scale=(9 * Ch) ** -0.5)
    bg = torch.randn(C, device=dev) * 0.1; bb = torch.randn(C, device=dev) * 0.1
    h = rnd(T, H, W, C); res = rnd(T, H, W, C)
    gw = torch.randn(C, device=dev); gb = torch.randn(C, device=dev)
    hf = h.float().permute(0, 3, 1, 2)
    g = hf.reshape(T, 32, -1)
    mean = g.mean(-1); var = g.var(-1, unbiased=False); rstd = (var + 1e-5).rsqrt()
    stats = torch.stack([mean, rstd], -1).contiguous()
    wp = ops.interleave_pair(ops.pack_conv_weight(wg), ops.pack_conv_weight(wb)); bp = ops.interleave_pair(bg, bb)
    out = ops.conv_gemm(actv, wp, taps=9, bias=bp, epilogue=ops.EPI_SPADE, h=h, gn_stats=stats, gn_weight=gw, gn_bias=gb, groups=32, res=res, beta=1.0)
    af = actv.float().permute(0, 3, 1, 2)
    gamma = F.conv2d(af, wg.float(), bg, padding=1); beta = F.conv2d(af, wb.float(), bb, padding=1)
    xn = F.group_norm(hf, 32, gw, gb, eps=1e-5)
    ref = (res.float().permute(0, 3, 1, 2) + xn * (1 + gamma) + beta).permute(0, 2, 3, 1)
    return report(f"spade T{T} {H}x{W} {Ch}->{C}", out, ref)

CASES = {
    "g1": lambda: case_gemm(256, 64, 128, bias=False),
    "g2": lambda: case_gemm(1024, 320, 320),
    "g3": lambda: case_gemm(20480, 320, 960, act=ops.ACT_NONE),
    "g4": lambda: case_gemm(1000, 1280, 1280, res=True),
    "g5": lambda: case_gemm(77, 1024, 640, bias=False),
    "g6": lambda: case_gemm(512, 128, 32),
    "g7": lambda: case_gemm(640, 512, 256, f32=True, act=ops.ACT_SILU),
    "g8": lambda: case_gemm(2048, 640, 1280, block_n=128, act=ops.ACT_GELU),
    "c1": lambda: case_conv3(2, 16, 16, 64, 64),
    "c2": lambda: case_conv3(5, 64, 64, 320, 320),
    "c3": lambda: case_conv3(5, 8, 8, 1280, 1280),
    "c4": lambda: case_conv3(5, 32, 32, 960, 640, two=True),
    "c5": lambda: case_conv3(3, 30, 46, 128, 256),
    "c6": lambda: case_conv3(1, 120, 120, 128, 128),
    "t1": lambda: case_t3(5, 8, 8, 1280),
    "t2": lambda: case_t3(5, 32, 32, 256),
    "e1": lambda: case_geglu(4096, 320, 1280),
    "e2": lambda: case_geglu(320, 1280, 5120),
    "s1": lambda: case_spade(5, 16, 16, 128, 1280),
    "s2": lambda: case_spade(5, 64, 64, 128, 320),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    bad = 0
    for n in names:
        try:
            ok = CASES[n]()
            torch.cuda.synchronize()
        except Exception as e:
            print(f"ERROR {n}: {type(e).__name__}: {e}", flush=True)
            ok = False
            if "CUDA error" in str(e) or "cuda" in str(e).lower():
                print("context likely dead; aborting this process", flush=True)
                sys.exit(2)
        bad += (not ok)
    print(f"done: {len(names) - bad}/{len(names)} pass")
    sys.exit(1 if bad else 0)
