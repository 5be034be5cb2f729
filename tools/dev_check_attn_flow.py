"""GPU dev check: attention + flow kernels vs torch.  Usage: python tools/dev_check_attn_flow.py [case ...]"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mgld_vsr_b200 import ops
torch.manual_seed(0)
dev = "cuda"

def report(name, got, ref, tol=4e-3):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs().max().item(); den = ref.abs().max().item() + 1e-12
    ok = err / den < tol and not math.isnan(err)
    print(f"{'PASS' if ok else 'FAIL'} {name}: max_abs_err={err:.4e} ref_max={den:.3e} rel={err/den:.3e}", flush=True)
    return ok

def attn_self(B, N, heads, dh, qscale=1.0):
    C = heads * dh
    qkv = (torch.randn(B * N, 3 * C, device=dev) * qscale).half()
    out = ops.attention(qkv, qkv, qkv, batch=B, heads=heads, head_dim=dh, nq=N, nkv=N, scale=dh ** -0.5,
                        q_col0=0, k_col0=C, v_col0=2 * C)
    q, k, v = [t.float().reshape(B, N, heads, dh).transpose(1, 2) for t in qkv.split(C, dim=1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, C)
    return report(f"self-attn B{B} N{N} h{heads} d{dh} qs{qscale}", out, ref)

def attn_cross(B, N, heads, dh, nkv=77):
    C = heads * dh
    q = torch.randn(B * N, C, device=dev).half()
    kv = torch.randn(nkv, 2 * C, device=dev).half()
    out = ops.attention(q, kv, kv, batch=B, heads=heads, head_dim=dh, nq=N, nkv=nkv, scale=dh ** -0.5,
                        k_col0=0, v_col0=C, kv_batched=False)
    qq = q.float().reshape(B, N, heads, dh).transpose(1, 2)
    k = kv[:, :C].float().reshape(1, nkv, heads, dh).transpose(1, 2).expand(B, -1, -1, -1)
    v = kv[:, C:].float().reshape(1, nkv, heads, dh).transpose(1, 2).expand(B, -1, -1, -1)
    ref = F.scaled_dot_product_attention(qq, k, v).transpose(1, 2).reshape(B * N, C)
    return report(f"cross-attn B{B} N{N} h{heads} d{dh} nkv{nkv}", out, ref)

def attn_legacy(B, N, heads, dh):
    # QKVAttentionLegacy layout: per head [q|k|v] interleaved: col = h*3*dh + {0,dh,2dh}
    C = heads * dh
    qkv = torch.randn(B * N, 3 * C, device=dev).half()
    out = ops.attention(qkv, qkv, qkv, batch=B, heads=heads, head_dim=dh, nq=N, nkv=N, scale=dh ** -0.5,
                        q_col0=0, k_col0=dh, v_col0=2 * dh, q_head_stride=3 * dh, k_head_stride=3 * dh, v_head_stride=3 * dh)
    x = qkv.float().reshape(B, N, heads, 3, dh)
    q, k, v = [x[:, :, :, i].transpose(1, 2) for i in range(3)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, C)
    return report(f"legacy-attn B{B} N{N} h{heads} d{dh}", out, ref)

def smooth_flow(n, h, w, mag=3.0):
    f = torch.randn(n, 2, 8, 8, device=dev) * mag
    return F.interpolate(f, size=(h, w), mode="bicubic", align_corners=False)

def flow_cases():
    ok = True
    for (n, c, h, w) in [(4, 4, 64, 64), (2, 3, 37, 53), (1, 2, 136, 240)]:
        x = torch.randn(n, c, h, w, device=dev); fl = smooth_flow(n, h, w)
        flp = fl.permute(0, 2, 3, 1).contiguous()
        grid_y, grid_x = torch.meshgrid(torch.arange(h, device=dev).float(), torch.arange(w, device=dev).float(), indexing="ij")
        vg = torch.stack((grid_x, grid_y), 2)[None] + flp
        vs = torch.stack((2.0 * vg[..., 0] / max(w - 1, 1) - 1.0, 2.0 * vg[..., 1] / max(h - 1, 1) - 1.0), dim=3)
        for border in (False, True):
            ref = F.grid_sample(x, vs, mode="bilinear", padding_mode="border" if border else "zeros", align_corners=True)
            ok &= report(f"flow_warp {n,c,h,w} border{border}", ops.flow_warp_f32(x, flp, 0, border=border), ref, 1e-5)
        ok &= report(f"flow_warp layout1 {n,c,h,w}", ops.flow_warp_f32(x, fl, 1), F.grid_sample(x, vs, align_corners=True), 1e-5)
        refn = F.grid_sample(x, vs, mode="nearest", align_corners=True)
        gotn = ops.flow_warp_f32(x, flp, 0, nearest=True)
        frac = ((gotn - refn).abs() > 1e-6).float().mean().item()
        print(f"{'PASS' if frac < 1e-3 else 'FAIL'} flow_warp nearest mismatch frac {frac:.2e}"); ok &= frac < 1e-3
        # adjoint
        xr = x.clone().requires_grad_(True)
        y = F.grid_sample(xr, vs, align_corners=True); g = torch.randn_like(y); y.backward(g)
        ok &= report(f"flow_warp_bwd {n,c,h,w}", ops.flow_warp_bwd_input_f32(g, flp, 0), xr.grad, 1e-5)
        # resize_flow
        oh, ow = h // 2 + 3, w // 2 + 1
        inp = fl.clone(); inp[:, 0] *= ow / w; inp[:, 1] *= oh / h
        ok &= report(f"resize_flow {h,w}->{oh,ow}", ops.resize_flow_f32(fl, oh, ow), F.interpolate(inp, size=(oh, ow), mode="bilinear", align_corners=False), 1e-5)
    return ok

def ref_fb(fwd_flow, bwd_flow, alpha=0.01, beta=0.5):
    def warp(feat, flow):
        b, c, h, w = feat.shape
        y, x = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
        grid = torch.stack([x, y], 0).float()[None] + flow
        xg = 2 * grid[:, 0] / (w - 1) - 1; yg = 2 * grid[:, 1] / (h - 1) - 1
        return F.grid_sample(feat, torch.stack([xg, yg], -1), mode="bilinear", padding_mode="zeros", align_corners=True)
    mag = torch.norm(fwd_flow, dim=1) + torch.norm(bwd_flow, dim=1)
    df = torch.norm(fwd_flow + warp(bwd_flow, fwd_flow), dim=1); db = torch.norm(bwd_flow + warp(fwd_flow, bwd_flow), dim=1)
    thr = alpha * mag + beta
    return (df > thr).float(), (db > thr).float()

def ref_guidance(z, ff, fb, fo, bo):
    # restatement of compute_temporal_condition_v4 (ddpm.py:3538-3574) with autograd
    def fw(x, flow):  # arch_util.flow_warp, flow (1,2,h,w) -> permuted
        _, _, h, w = x.shape
        fl = flow.permute(0, 2, 3, 1)
        gy, gx = torch.meshgrid(torch.arange(h, device=dev).float(), torch.arange(w, device=dev).float(), indexing="ij")
        vg = torch.stack((gx, gy), 2)[None] + fl
        vs = torch.stack((2.0 * vg[..., 0] / max(w - 1, 1) - 1.0, 2.0 * vg[..., 1] / max(h - 1, 1) - 1.0), dim=3)
        return F.grid_sample(x, vs, mode="bilinear", padding_mode="zeros", align_corners=True)
    t = z.shape[0]
    lat = z[None]  # b=1
    loss_b = 0; warp = torch.zeros_like(lat[:, -1])
    for i in range(t - 1, -1, -1):
        cur = lat[:, i]
        if i < t - 1:
            warp = fw(cur, fb[None, i]); m = 1 - fo[None, i, None]
            loss_b = loss_b + F.l1_loss(m * prev, m * cur)
        prev = warp
    loss_f = 0; warp = torch.zeros_like(lat[:, 0])
    for i in range(t):
        cur = lat[:, i]
        if i > 0:
            warp = fw(cur, ff[None, i - 1]); m = 1 - bo[None, i - 1, None]
            loss_f = loss_f + F.l1_loss(m * prev, m * cur)
        prev = warp
    return loss_b + loss_f

def guidance_cases():
    ok = True
    for (t, c, h, w) in [(5, 4, 64, 64), (5, 4, 136, 240), (2, 4, 32, 48), (3, 4, 20, 28)]:
        z = torch.randn(t, c, h, w, device=dev)
        ff = smooth_flow(t - 1, h, w, 1.5); fb = -ff + 0.3 * smooth_flow(t - 1, h, w, 1.0)
        fo, bo = ops.fb_consistency_f32(fb, ff)   # script arg order: fwd=flows[1], bwd=flows[0]
        rfo, rbo = ref_fb(fb, ff)
        mism = ((fo != rfo).float().mean() + (bo != rbo).float().mean()).item()
        print(f"{'PASS' if mism < 1e-4 else 'FAIL'} fb_consistency {t,h,w}: mismatch frac {mism:.2e} occ_frac {fo.mean().item():.3f}"); ok &= mism < 1e-4
        zr = z.clone().requires_grad_(True)
        loss = ref_guidance(zr, ff, fb, rfo, rbo)
        g = torch.autograd.grad(loss, zr)[0]
        step = -10.0 * -3.2
        ref = z - step * g
        out, l, gws = ops.motion_guidance_f32(z, ff, fb, rfo, rbo, step, want_loss=True)
        ok &= report(f"guidance loss {t,c,h,w}", l, loss.detach().reshape(1), 1e-5)
        ok &= report(f"guidance grad {t,c,h,w}", gws, g, 1e-4)
        ok &= report(f"guidance step {t,c,h,w}", out, ref, 1e-5)
    return ok

CASES = {
    "a1": lambda: attn_self(1, 128, 1, 64),
    "a2": lambda: attn_self(2, 256, 5, 64),
    "a3": lambda: attn_self(5, 4096, 5, 64),
    "a4": lambda: attn_self(5, 1024, 10, 64, qscale=3.0),
    "a5": lambda: attn_self(5, 64, 20, 64),
    "a6": lambda: attn_self(2, 200, 3, 64),
    "x1": lambda: attn_cross(5, 4096, 5, 64),
    "x2": lambda: attn_cross(5, 64, 20, 64),
    "l1": lambda: attn_legacy(5, 4096, 4, 64),
    "l2": lambda: attn_legacy(5, 256, 4, 128),
    "l3": lambda: attn_legacy(5, 64, 4, 128),
    "fl": flow_cases,
    "gd": guidance_cases,
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    bad = 0
    for n in names:
        try:
            ok = CASES[n](); torch.cuda.synchronize()
        except Exception as e:
            import traceback; traceback.print_exc()
            print(f"ERROR {n}: {type(e).__name__}: {e}", flush=True); ok = False
            if "CUDA" in str(e) or "cuda" in str(e): sys.exit(2)
        bad += (not ok)
    print(f"done: {len(names) - bad}/{len(names)} pass"); sys.exit(1 if bad else 0)
