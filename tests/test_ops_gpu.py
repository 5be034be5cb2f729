"""GPU parity of every C-ABI op against a plain PyTorch fp32 evaluation of the same op (tests/emu_ops.py), on
fp16-rounded seeded inputs.  Tolerances: fp16-output ops 4e-3 of the output range (one fp16 rounding + fp32
accumulation-order noise); fp32 ops 1e-4."""
import pytest
import torch
import torch.nn.functional as F

import emu_ops as E
from common import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ops():
    from mgld_vsr_b200 import ops as O
    return O


def rnd(*s, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(s))
    return (torch.randn(*s, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("M,K,N,kw", [
    (256, 64, 128, {}), (1024, 320, 320, {}), (20480, 320, 960, {}), (77, 1024, 640, {}), (512, 128, 32, {}),
    (1000, 1280, 1280, dict(res=True)), (640, 512, 256, dict(f32=True, act=2)), (2048, 640, 1280, dict(act=4, bn=128)),
    (300, 200, 512, {}),      # K tail (200 = 3*64 + 8): TMA zero fill
])
def test_gemm(M, K, N, kw):
    O = ops()
    a, w = rnd(M, K).half(), rnd(N, K, scale=K ** -0.5).half()
    b = rnd(N)
    r = rnd(M, N).half() if kw.get("res") else None
    args = dict(bias=b, act=kw.get("act", 0), res=r, alpha=0.5 if r is not None else 1.0, beta=2.0 if r is not None else 0.0,
                out_f32=kw.get("f32", False))
    got = O.conv_gemm(a, w, block_n=kw.get("bn", 0), **args)
    ref = E.conv_gemm(a, w, **args)
    assert rel_err(got, ref) < 4e-3


@pytest.mark.parametrize("T,H,W,Ci,Co,two", [(2, 16, 16, 64, 64, False), (5, 64, 64, 320, 320, False),
                                            (5, 8, 8, 1280, 1280, False), (5, 32, 32, 960, 640, True),
                                            (3, 30, 46, 128, 256, False), (1, 120, 120, 128, 128, False)])
def test_conv3x3(T, H, W, Ci, Co, two):
    O = ops()
    x = rnd(T, H, W, Ci).half()
    w = O.pack_conv_weight(rnd(Co, Ci, 3, 3, scale=(9 * Ci) ** -0.5))
    b = rnd(Co)
    if two:
        c1 = (Ci // 128) * 64
        a, a2 = x[..., :c1].contiguous(), x[..., c1:].contiguous()
    else:
        a, a2 = x, None
    got = O.conv_gemm(a, w, taps=9, a2=a2, bias=b)
    ref = E.conv_gemm(a, w, taps=9, a2=a2, bias=b)
    assert rel_err(got, ref) < 4e-3


def test_conv_strided_views_and_column_offset():
    """the RDB usage: A is a channel-prefix view of a wider buffer, the output lands in a column slot of it"""
    O = ops()
    T, H, W, C = 2, 16, 24, 128
    buf = torch.zeros(T, H, W, C + 256, device=DEV, dtype=torch.float16)
    buf[..., :C + 64] = rnd(T, H, W, C + 64).half()
    w = O.pack_conv_weight(rnd(32, C + 64, 3, 3, scale=0.03))
    b = rnd(32)
    ref_buf = buf.clone()
    O.conv_gemm(buf[..., :C + 64], w, taps=9, bias=b, act=O.ACT_LRELU02, out=buf, out_col0=C + 64, block_n=32)
    E.conv_gemm(ref_buf[..., :C + 64], w, taps=9, bias=b, act=E.ACT_LRELU02, out=ref_buf, out_col0=C + 64)
    assert rel_err(buf, ref_buf) < 4e-3
    assert torch.equal(buf[..., C + 96:], ref_buf[..., C + 96:])     # untouched slots stay zero


def test_split_k_path():
    """few output tiles + long K -> partial fp32 tiles + deterministic finalize; must equal the single-pass result"""
    O = ops()
    T, H, W, C = 5, 8, 8, 1280
    x, res = rnd(T, H, W, C).half(), rnd(T, H, W, C).half()
    w, b = O.pack_conv_weight(rnd(C, C, 3, 3, scale=(9 * C) ** -0.5)), rnd(C)
    kw = dict(taps=9, bias=b, act=O.ACT_SILU, res=res, alpha=0.5, beta=2.0)
    O.SPLIT_K = False
    single = O.conv_gemm(x, w, **kw)
    O.SPLIT_K = True
    n0 = O.LAUNCHES[0]
    split = O.conv_gemm(x, w, **kw)
    assert O.LAUNCHES[0] - n0 == 2, "this shape is expected to take the split-K path (2 launches)"
    assert rel_err(split, E.conv_gemm(x, w, **kw)) < 4e-3 and rel_err(split, single) < 2e-3
    again = O.conv_gemm(x, w, **kw)
    assert torch.equal(split, again)                              # fixed summation order: bitwise reproducible
    # pair epilogues through the finalize kernel
    a = rnd(320, 1280).half()
    wf, bf = rnd(1280, 1280, scale=1280 ** -0.5).half(), rnd(1280)
    wp, bp = O.interleave_pair(wf[:640], wf[640:]), O.interleave_pair(bf[:640], bf[640:])
    y = a.float() @ wf.float().t() + bf
    assert rel_err(O.conv_gemm(a, wp, bias=bp, epilogue=O.EPI_GEGLU), y[:, :640] * F.gelu(y[:, 640:])) < 4e-3
    actv, h = rnd(T, H, W, 128).half(), rnd(T, H, W, 640).half()
    wg, wb = O.pack_conv_weight(rnd(640, 128, 3, 3, scale=0.03)), O.pack_conv_weight(rnd(640, 128, 3, 3, scale=0.03, seed=5))
    st = O.gn_finalize(O.gn_stats(h), H * W, 640, 1e-5)
    kw = dict(taps=9, bias=O.interleave_pair(rnd(640, scale=0.1), rnd(640, scale=0.1, seed=3)), epilogue=O.EPI_SPADE, h=h,
              gn_stats=st, gn_weight=rnd(640, seed=7), gn_bias=rnd(640, seed=9), groups=32, res=rnd(T, H, W, 640).half(), beta=1.0)
    wp = O.interleave_pair(wg, wb)
    assert rel_err(O.conv_gemm(actv, wp, **kw), E.conv_gemm(actv, wp, **kw)) < 4e-3


@pytest.mark.parametrize("T,H,W,Ci,Co,taps,res", [(5, 64, 64, 320, 320, 9, False), (5, 8, 8, 1280, 1280, 9, True),
                                                  (5, 1, 1024, 640, 640, 1, True), (3, 30, 46, 128, 256, 9, False),
                                                  (5, 16, 16, 256, 128, 9, False)])
def test_fused_groupnorm_statistics(T, H, W, Ci, Co, taps, res):
    """the (sum, sumsq) a conv epilogue accumulates for its consumer's GroupNorm == gn_stats of the stored output"""
    O = ops()
    x = rnd(T, H, W, Ci).half()
    if H == 1:
        x = x.reshape(T, W, Ci)
    w = (rnd(Co, taps * Ci, scale=(taps * Ci) ** -0.5)).half()
    r = rnd(*x.shape[:-1], Co).half() if res else None
    sums = torch.zeros(T, 32, 2, 2, device=DEV, dtype=torch.float64)     # 16-byte fixed-point accumulators (mgld.h)
    out = O.conv_gemm(x, w, taps=taps, bias=rnd(Co), res=r, beta=1.0 if res else 0.0, stats_out=sums)
    ref = O.gn_stats(out.reshape(T, -1, Co))
    assert torch.equal(sums.view(torch.int64), ref.view(torch.int64))    # order-independent accumulation: same bits


@pytest.mark.parametrize("cta_pair", ["0", "1"])
@pytest.mark.parametrize("T,H,W,Ci,Co,groups", [(5, 64, 64, 320, 320, 32), (10, 64, 64, 320, 320, 32), (2, 32, 32, 640, 640, 32),
                                               (5, 64, 64, 256, 256, 32), (2, 128, 128, 128, 128, 32), (2, 64, 64, 960, 320, 32),
                                               (3, 30, 46, 128, 256, 32), (5, 16, 16, 1280, 1280, 32), (1, 64, 64, 320, 320, 8)])
def test_conv_epilogue_groupnorm_statistics(monkeypatch, cta_pair, T, H, W, Ci, Co, groups):
    """MGLD_CONV_FUSED_STATS=1: a long-K 3x3 conv accumulates the (sum, sumsq) of its output for the consumer's GroupNorm in
    its epilogue (conv_gemm.cu variant 7: column sums of every finished staging panel) instead of a streaming pass over the
    stored output.  Same values up to the fp32 rounding of the per-thread partial sums; the output itself is bit-identical.
    Shapes that do not qualify (partial tiles, split-K) fall back to the streaming pass inside the same call."""
    O = ops()
    x = rnd(T, H, W, Ci).half()
    w = (rnd(Co, 9 * Ci, scale=(9 * Ci) ** -0.5)).half()
    b = rnd(Co)
    monkeypatch.setenv("MGLD_CONV_PAIR", cta_pair)
    monkeypatch.setenv("MGLD_CONV_FUSED_STATS", "0")
    ref_out = O.conv_gemm(x, w, taps=9, bias=b)
    ref = O.gn_stats(ref_out.reshape(T, -1, Co), groups=groups)
    for rep in range(2):                                                  # twice: bitwise repeatable
        monkeypatch.setenv("MGLD_CONV_FUSED_STATS", "1")
        sums = torch.zeros(T, groups, 2, 2, device=DEV, dtype=torch.float64)
        out = O.conv_gemm(x, w, taps=9, bias=b, stats_out=sums, stats_groups=groups)
        torch.cuda.synchronize()
        assert torch.equal(out, ref_out)
        got, want = O.decode_sums(sums), O.decode_sums(ref)
        assert torch.allclose(got, want, rtol=2e-5, atol=1e-3), (got - want).abs().max()
        if rep == 0:
            first = sums.clone()
            print(f"[fused stats pair={cta_pair} {T}x{H}x{W} {Ci}->{Co}] same bits as the streaming pass: "
                  f"{torch.equal(sums.view(torch.int64), ref.view(torch.int64))}, max rel diff "
                  f"{((got - want).abs() / want.abs().clamp_min(1.0)).max():.1e}")
        else:
            assert torch.equal(sums.view(torch.int64), first.view(torch.int64))


@pytest.mark.parametrize("T,H,W,C", [(5, 8, 8, 1280), (5, 32, 32, 256), (2, 12, 20, 128)])
def test_temporal_conv(T, H, W, C):
    O = ops()
    x = rnd(T, H, W, C).half()
    w = O.pack_temporal_weight(rnd(C, C, 3, 1, 1, scale=(3 * C) ** -0.5))
    b = rnd(C)
    kw = dict(taps=3, bias=b, alpha=0.3, beta=0.7, res=x)
    assert rel_err(O.conv_gemm(x, w, **kw), E.conv_gemm(x, w, **kw)) < 4e-3


@pytest.mark.parametrize("cta_pair", ["0", "1"])
def test_pair_epilogue_tile_widths_and_unit_order(monkeypatch, cta_pair):
    """GEGLU / SPADE with one (block_n 128) and two (block_n 256) output panels per tile, N-major and M-major order of the
    work units, single CTAs and CTA pairs: the same accumulation order per output element, so the same bits."""
    O = ops()
    monkeypatch.setenv("MGLD_CONV_PAIR", cta_pair)
    T, H, W = 3, 20, 28                      # ragged pixel tiles
    a = rnd(T, H, W, 320).half()
    w, b = rnd(1536, 320, scale=320 ** -0.5).half(), rnd(1536)
    wp, bp = O.interleave_pair(w[:768], w[768:]), O.interleave_pair(b[:768], b[768:])
    y = a.float() @ w.float().t() + b
    ref_g = y[..., :768] * F.gelu(y[..., 768:])
    Ch, C = 128, 256
    actv, h, res = rnd(T, H, W, Ch).half(), rnd(T, H, W, C).half(), rnd(T, H, W, C).half()
    wg, wb = O.pack_conv_weight(rnd(C, Ch, 3, 3, scale=0.03)), O.pack_conv_weight(rnd(C, Ch, 3, 3, scale=0.03, seed=5))
    st = O.gn_finalize(O.gn_stats(h), H * W, C, 1e-5)
    kw = dict(taps=9, bias=O.interleave_pair(rnd(C, scale=0.1), rnd(C, scale=0.1, seed=3)), epilogue=O.EPI_SPADE, h=h, gn_stats=st,
              gn_weight=rnd(C, seed=7), gn_bias=rnd(C, seed=9), groups=32, res=res, beta=1.0)
    ws = O.interleave_pair(wg, wb)
    ref_s = E.conv_gemm(actv, ws, **kw)
    outs = []
    for bn in ("128", "256"):
        for raster in ("0", "1"):
            monkeypatch.setenv("MGLD_CONV_PAIR_EPI_BN", bn)
            monkeypatch.setenv("MGLD_CONV_RASTER", raster)
            g = O.conv_gemm(a, wp, bias=bp, epilogue=O.EPI_GEGLU)
            sp = O.conv_gemm(actv, ws, **kw)
            lin = O.conv_gemm(a, w, bias=b, res=None)
            assert rel_err(g, ref_g) < 4e-3 and rel_err(sp, ref_s) < 4e-3 and rel_err(lin, y) < 4e-3, (bn, raster)
            outs.append((g, sp, lin))
    for o in outs[1:]:
        assert all(torch.equal(x, y0) for x, y0 in zip(o, outs[0]))


def test_geglu_and_spade_epilogues():
    O = ops()
    a = rnd(4096, 320).half()
    w, b = rnd(2560, 320, scale=320 ** -0.5).half(), rnd(2560)
    wp, bp = O.interleave_pair(w[:1280], w[1280:]), O.interleave_pair(b[:1280], b[1280:])
    got = O.conv_gemm(a, wp, bias=bp, epilogue=O.EPI_GEGLU)
    y = a.float() @ w.float().t() + b
    assert rel_err(got, y[:, :1280] * F.gelu(y[:, 1280:])) < 4e-3
    T, H, W, Ch, C = 5, 16, 16, 128, 640
    actv, h, res = rnd(T, H, W, Ch).half(), rnd(T, H, W, C).half(), rnd(T, H, W, C).half()
    wg, wb = O.pack_conv_weight(rnd(C, Ch, 3, 3, scale=0.03)), O.pack_conv_weight(rnd(C, Ch, 3, 3, scale=0.03, seed=5))
    bg, bb, gw, gb = rnd(C, scale=0.1), rnd(C, scale=0.1, seed=3), rnd(C, seed=7), rnd(C, seed=9)
    st = O.gn_finalize(O.gn_stats(h), H * W, C, 1e-5)
    kw = dict(taps=9, bias=O.interleave_pair(bg, bb), epilogue=O.EPI_SPADE, h=h, gn_stats=st, gn_weight=gw, gn_bias=gb,
              groups=32, res=res, beta=1.0)
    wp = O.interleave_pair(wg, wb)
    assert rel_err(O.conv_gemm(actv, wp, **kw), E.conv_gemm(actv, wp, **kw)) < 4e-3


@pytest.mark.parametrize("B,N,heads,dh,mode", [(1, 128, 1, 64, "self"), (5, 4096, 5, 64, "self"), (5, 1024, 10, 64, "self"),
                                              (5, 64, 20, 64, "self"), (2, 200, 3, 64, "self"), (2, 300, 3, 64, "self"),
                                              (1, 520, 2, 64, "self"), (5, 4096, 5, 64, "cross"),
                                              (5, 64, 20, 64, "cross"), (5, 4096, 4, 64, "legacy"), (5, 256, 4, 128, "legacy"),
                                              (5, 64, 4, 128, "legacy")])
def test_attention(B, N, heads, dh, mode):
    O = ops()
    C = heads * dh
    if mode == "cross":
        q, kv = rnd(B * N, C).half(), rnd(77, 2 * C).half()
        kw = dict(batch=B, heads=heads, head_dim=dh, nq=N, nkv=77, scale=dh ** -0.5, k_col0=0, v_col0=C, kv_batched=False)
        got, ref = O.attention(q, kv, kv, **kw), E.attention(q.cpu(), kv.cpu(), kv.cpu(), **kw)
    else:
        qkv = rnd(B * N, 3 * C).half()
        if mode == "self":
            kw = dict(q_col0=0, k_col0=C, v_col0=2 * C)
        else:
            kw = dict(q_col0=0, k_col0=dh, v_col0=2 * dh, q_head_stride=3 * dh, k_head_stride=3 * dh, v_head_stride=3 * dh)
        kw.update(batch=B, heads=heads, head_dim=dh, nq=N, nkv=N, scale=dh ** -0.5)
        got = O.attention(qkv, qkv, qkv, **kw)
        x = qkv.float().reshape(B, N, 3, heads, dh) if mode == "self" else qkv.float().reshape(B, N, heads, 3, dh).transpose(2, 3)
        q, k, v = [x[:, :, i].transpose(1, 2) for i in range(3)]
        ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, C)
    assert rel_err(got.cpu(), ref.cpu()) < 4e-3


def _sdpa_ref(qkv, B, N, heads, dh):
    x = qkv.float().reshape(B, N, 3, heads, dh)
    q, k, v = [x[:, :, i].transpose(1, 2) for i in range(3)]
    return F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, heads * dh)


@pytest.mark.parametrize("case", ["later_blocks_overflow", "peaky_single_key", "rising_scores", "first_block_dominates"])
@pytest.mark.parametrize("B,N,heads", [(2, 1024, 2), (1, 4096, 1), (1, 600, 1)])
def test_attention_no_max_fast_path_fallback(case, B, N, heads):
    """attention_v3 skips the row max after key block 0 and redoes a block when its row sum reaches 2^15
    (attention.cu `exp_store` / `over`).  N(0,1) logits never get there; these inputs do: later key blocks whose scores
    exceed block 0's maximum by far more than 2^15 in the exp2 domain, one dominating key (peaky softmax, as real
    checkpoints produce), scores that keep rising block after block (repeated fallbacks), and the opposite case where
    block 0 dominates (later probabilities underflow)."""
    O = ops()
    dh, C = 64, heads * 64
    qkv = rnd(B * N, 3 * C, seed=11).reshape(B, N, 3, heads, dh)
    if case == "later_blocks_overflow":
        qkv[:, :128, 1] *= 0.1                      # block 0: logits ~ N(0, 0.1^2)
        qkv[:, 128:, 1] *= 8.0                      # later blocks: logits ~ N(0, 8^2) -> max ~ +25 nats >> 15 * ln 2
    elif case == "peaky_single_key":
        qkv[:, N // 2 + 3, 1] = qkv[:, :, 0].mean(1) * 40.0 + 6.0   # one key far above all others for most queries
    elif case == "rising_scores":
        ramp = torch.linspace(0.2, 10.0, N, device=DEV).reshape(1, N, 1, 1)
        qkv[:, :, 1] *= ramp
    else:
        qkv[:, :128, 1] *= 10.0
        qkv[:, 128:, 1] *= 0.05
    qkv = qkv.reshape(B * N, 3 * C).half()
    kw = dict(batch=B, heads=heads, head_dim=dh, nq=N, nkv=N, scale=dh ** -0.5, q_col0=0, k_col0=C, v_col0=2 * C)
    got = O.attention(qkv, qkv, qkv, **kw)
    ref = _sdpa_ref(qkv, B, N, heads, dh)
    assert torch.isfinite(got).all()
    assert rel_err(got, ref) < 4e-3, rel_err(got, ref)


@pytest.mark.parametrize("group", ["4", "2"])
@pytest.mark.parametrize("B,N,gain", [(1, 64, 1.0), (2, 200, 1.0), (2, 4096, 1.0), (1, 14400, 1.0), (1, 1000, 6.0), (1, 1024, -1.0)])
def test_attention_head_dim_512(monkeypatch, group, B, N, gain):
    """the VAEs' single-head middle attention (model.py:247-305) on the split-D flash kernel (attention_hd512.cu): one
    key block, ragged tails of the 128-row query tile and the 64-key block, a 512^2 frame (N = 4096), a 960^2 VAE tile
    (N = 14400), peaky logits, and scores that keep rising with the key index (every block moves the lazy maximum: the
    O rescale path); with 4 and 2 slabs per TMA operation."""
    if group != "4" and N > 4096:
        pytest.skip("the large case runs once")
    monkeypatch.setenv("MGLD_HD512_GROUP", group)
    O = ops()
    C = 512
    qkv = rnd(B * N, 3 * C, seed=5).reshape(B, N, 3, C)
    if gain > 0:
        qkv[:, :, 1] *= gain
    else:
        qkv[:, :, 1] *= torch.linspace(0.2, 8.0, N, device=DEV).reshape(1, N, 1)
    qkv = qkv.reshape(B * N, 3 * C).half()
    kw = dict(batch=B, heads=1, head_dim=C, nq=N, nkv=N, scale=C ** -0.5, q_col0=0, k_col0=C, v_col0=2 * C)
    got = O.attention(qkv, qkv, qkv, **kw)
    ref = _sdpa_ref(qkv, B, N, 1, C)
    assert torch.isfinite(got).all()
    err = rel_err(got, ref)
    print(f"[hd512 group={group} B={B} N={N} gain={gain}] rel err {err:.2e}; per 64-col slab: "
          + " ".join(f"{rel_err(got[:, c:c + 64], ref[:, c:c + 64]):.1e}" for c in range(0, C, 64)))
    assert err < 4e-3, err


def test_attention_cross_77_keys_full_shape():
    """the 77-key cross-attention variant at the shapes the UNet runs it (15 of the 34 attention launches of a
    tile-step): N=4096 x 5 heads, 1024 x 10, 256 x 20, K/V broadcast over the frames, peaky logits included"""
    O = ops()
    for (B, N, heads, gain) in [(5, 4096, 5, 1.0), (5, 1024, 10, 1.0), (5, 256, 20, 1.0), (5, 4096, 5, 6.0)]:
        C = heads * 64
        q, kv = rnd(B * N, C, seed=3).half(), (rnd(77, 2 * C, seed=4) * gain).half()
        kw = dict(batch=B, heads=heads, head_dim=64, nq=N, nkv=77, scale=0.125, k_col0=0, v_col0=C, kv_batched=False)
        got = O.attention(q, kv, kv, **kw)
        qh = q.float().reshape(B, N, heads, 64).transpose(1, 2)
        k = kv[:, :C].float().reshape(1, 77, heads, 64).transpose(1, 2).expand(B, -1, -1, -1)
        v = kv[:, C:].float().reshape(1, 77, heads, 64).transpose(1, 2).expand(B, -1, -1, -1)
        ref = F.scaled_dot_product_attention(qh, k, v).transpose(1, 2).reshape(B * N, C)
        assert rel_err(got, ref) < 4e-3, (N, heads, gain, rel_err(got, ref))


@pytest.mark.parametrize("T,HW,C1,C2", [(5, 4096, 320, 0), (5, 1024, 1280, 640), (2, 64, 64, 0), (3, 900, 128, 128),
                                       (5, 256, 2560, 0)])
def test_groupnorm(T, HW, C1, C2):
    O = ops()
    x1 = (rnd(T, HW, C1) * 2 + 0.5).half()
    x2 = rnd(T, HW, C2).half() if C2 else None
    g, b = rnd(C1 + C2), rnd(C1 + C2, seed=2)
    sums = O.gn_stats(x1, x2)
    dsum = O.decode_sums(sums)                                             # float64 (sum, sumsq) [T, G, 2]
    assert rel_err(dsum, E.gn_stats(x1, x2)) < 1e-5
    for _ in range(3):                                                     # fixed-point accumulation: bitwise repeatable
        assert torch.equal(O.gn_stats(x1, x2).view(torch.int64), sums.view(torch.int64))
    assert rel_err(O.gn_apply(x1, sums, 1e-5, g, b, True, x2=x2), E.gn_apply(x1, dsum, 1e-5, g, b, True, x2=x2)) < 3e-3
    assert rel_err(O.gn_apply(x1, sums, 1e-6, None, None, False, x2=x2), E.gn_apply(x1, dsum, 1e-6, None, None, False, x2=x2)) < 3e-3
    assert rel_err(O.gn_finalize(sums, HW, C1 + C2, 1e-5), E.gn_finalize(dsum, HW, C1 + C2, 1e-5)) < 1e-5


@pytest.mark.parametrize("T,HW,C1,C2", [(5, 4096, 320, 0), (5, 4096, 640, 320), (5, 1024, 1280, 640), (2, 64, 64, 0),
                                       (3, 900, 128, 128), (5, 256, 2560, 0), (1, 64, 1280, 0), (5, 4096, 320, 320),
                                       (1, 65536, 128, 0)])   # the last one exceeds the register patch: multi-pass path
def test_group_norm_single_launch(T, HW, C1, C2):
    """mgld_group_norm_f16 (cluster / DSMEM reduction, data kept in registers) against the torch emulation"""
    O = ops()
    x1 = (rnd(T, HW, C1) * 2 + 0.5).half()
    x2 = rnd(T, HW, C2).half() if C2 else None
    g, b = rnd(C1 + C2), rnd(C1 + C2, seed=2)
    got, st = O.group_norm(x1, g, b, 1e-5, True, x2=x2, want_stats=True)
    ref, rst = E.group_norm(x1, g, b, 1e-5, True, x2=x2, want_stats=True)
    assert rel_err(got, ref) < 3e-3
    assert rel_err(st, rst) < 1e-5
    got2 = O.group_norm(x1, None, None, 1e-6, False, x2=x2)
    assert rel_err(got2, E.group_norm(x1, None, None, 1e-6, False, x2=x2)) < 3e-3
    st2 = O.group_norm(x1, None, None, 1e-5, False, x2=x2, want_out=False, want_stats=True)
    assert rel_err(st2, rst) < 1e-5
    if O._L.lib().mgld_group_norm_fused_supported(C1 + C2, T, HW, 32):   # fixed-order reduction tree: bitwise repeatable
        assert torch.equal(got, O.group_norm(x1, g, b, 1e-5, True, x2=x2))


@pytest.mark.parametrize("M,C", [(20480, 320), (5120, 640), (333, 1280), (7, 64), (65, 256), (19, 1024), (11, 2048), (9, 1536)])
def test_layernorm(M, C):
    O = ops()
    x, g, b = (rnd(M, C) * 3 + 1).half(), rnd(C), rnd(C, seed=4)
    assert rel_err(O.layernorm(x, g, b), E.layernorm(x, g, b)) < 3e-3


def test_softmax_rows_and_small_ops():
    O = ops()
    s = rnd(300, 1000) * 20
    assert rel_err(O.softmax_rows(s, 0.044), E.softmax_rows(s, 0.044)) < 2e-3
    x = rnd(3, 5, 17, 23)
    assert torch.equal(O.nchw_to_nhwc(x), E.nchw_to_nhwc(x))
    xh = rnd(2, 9, 13, 64).half()
    assert torch.equal(O.nhwc_to_nchw(xh), E.nhwc_to_nchw(xh))
    assert torch.equal(O.upsample2x(xh), E.upsample2x(xh))
    for pad in (0, 1):
        assert torch.equal(O.im2col_s2(xh, pad), E.im2col_s2(xh, pad))
    xe = rnd(2, 10, 14, 64).half()
    for pad in (0, 1):
        assert torch.equal(O.im2col_s2(xe, pad), E.im2col_s2(xe, pad))
    y = rnd(2, 9, 13, 64, seed=8).half()
    assert rel_err(O.axpby(xh, y, 1.0, 0.7), E.axpby(xh, y, 1.0, 0.7)) < 2e-3


def test_stem_and_head_convs():
    O = ops()
    x = rnd(2, 4, 20, 28)
    w, b = rnd(64, 4, 3, 3, scale=0.2), rnd(64)
    assert rel_err(O.conv_small_cin(x, w, b), E.conv_small_cin(x, w, b)) < 2e-3
    # the three instantiations (4x3x3, 3x3x3, run-time tap count), pixel counts that are not multiples of the 64-pixel
    # batch or of the 4-pixel step, channel counts around the 128-channel chunk
    for (n, ci, h, wd, co, ks) in [(3, 4, 13, 11, 320, 3), (1, 3, 17, 9, 128, 3), (2, 8, 7, 5, 200, 3), (2, 4, 9, 7, 64, 1),
                                   (1, 1, 5, 5, 32, 3)]:
        x = rnd(n, ci, h, wd, seed=n + ci)
        w, b = rnd(co, ci, ks, ks, scale=0.2, seed=co), rnd(co, seed=ks)
        assert rel_err(O.conv_small_cin(x, w, b), E.conv_small_cin(x, w, b)) < 2e-3, (n, ci, h, wd, co, ks)
    w1, b1 = rnd(8, 8, 1, 1, scale=0.3), rnd(8)
    x8 = rnd(2, 8, 9, 11)
    assert rel_err(O.conv_small_f32(x8, w1, b1), E.conv_small_f32(x8, w1, b1)) < 1e-5
    xh = rnd(2, 12, 10, 320).half()
    for co in (3, 4, 8):
        wp, bb = O.pack_conv_weight(rnd(co, 320, 3, 3, scale=0.02)), rnd(co)
        assert rel_err(O.conv3x3_small_cout(xh, wp, bb), E.conv3x3_small_cout(xh, wp, bb)) < 1e-4


def test_time_embedding_gemv_temporal_attention_gaussian():
    O = ops()
    t = torch.tensor([937.0], device=DEV)
    assert rel_err(O.timestep_embedding(t, 320), E.timestep_embedding(t.cpu(), 320).to(DEV)) < 2e-4   # fp32 sin/cos of ~1e3 rad
    x, w, b, a = rnd(1280), rnd(640, 1280, scale=0.03).half(), rnd(640), rnd(640, seed=3)
    assert rel_err(O.gemv(x, w, b, a, True, False), E.gemv(x, w, b, a, True, False)) < 1e-4
    assert rel_err(O.gemv(x, w, b, None, False, True), E.gemv(x, w, b, None, False, True)) < 1e-4
    qkv = rnd(5, 64, 3 * 1280).half()
    assert rel_err(O.temporal_attention(qkv, 20, 0.125), E.temporal_attention(qkv, 20, 0.125)) < 3e-3
    m, n = rnd(2, 8, 6, 7), rnd(2, 4, 6, 7, seed=1)
    assert rel_err(O.gaussian_sample(m, n, 0.18215), E.gaussian_sample(m, n, 0.18215)) < 1e-6


def test_flow_ops():
    O = ops()
    for (n, c, h, w) in [(4, 4, 64, 64), (2, 3, 37, 53), (1, 2, 136, 240)]:
        x = rnd(n, c, h, w)
        fl = F.interpolate(rnd(n, 2, 8, 8) * 3, size=(h, w), mode="bicubic")
        flp = fl.permute(0, 2, 3, 1).contiguous()
        for border in (False, True):
            assert rel_err(O.flow_warp_f32(x, flp, 0, border=border), E.flow_warp_f32(x.cpu(), flp.cpu(), 0, border=border).to(DEV)) < 1e-4
        assert rel_err(O.flow_warp_f32(x, fl, 1), E.flow_warp_f32(x.cpu(), fl.cpu(), 1).to(DEV)) < 1e-4
        oh, ow = h // 2 + 3, w // 2 + 1
        assert rel_err(O.resize_flow_f32(fl, oh, ow), E.resize_flow_f32(fl, oh, ow)) < 1e-5


@pytest.mark.parametrize("t,c,h,w", [(5, 4, 64, 64), (5, 4, 136, 240), (2, 4, 32, 48), (3, 4, 20, 28), (1, 4, 16, 16)])
def test_guidance(t, c, h, w):
    O = ops()
    z = rnd(t, c, h, w)
    if t == 1:
        out = O.motion_guidance_f32(z, None, None, None, None, 3.0)
        assert torch.equal(out, z)
        return
    ff = F.interpolate(rnd(t - 1, 2, 8, 8) * 1.5, size=(h, w), mode="bicubic")
    fb = -ff + 0.3 * F.interpolate(rnd(t - 1, 2, 8, 8, seed=3), size=(h, w), mode="bicubic")
    fo, bo = O.fb_consistency_f32(fb, ff)
    rfo, rbo = E.fb_consistency_f32(fb.cpu(), ff.cpu())
    assert ((fo.cpu() != rfo).float().mean() + (bo.cpu() != rbo).float().mean()).item() < 1e-4
    assert 0.02 < fo.mean().item() < 0.98                      # masks are genuinely mixed
    out, loss, g = O.motion_guidance_f32(z, ff, fb, fo, bo, 32.0, want_loss=True)
    rout, rloss, rg = E.motion_guidance_f32(z.cpu(), ff.cpu(), fb.cpu(), fo.cpu(), bo.cpu(), 32.0, want_loss=True)
    assert rel_err(loss.cpu(), rloss) < 1e-5 and rel_err(g.cpu(), rg) < 1e-4 and rel_err(out.cpu(), rout) < 1e-5
    # the scatter-add of the warp adjoint accumulates in 64-bit fixed point: bitwise repeatable run to run
    for _ in range(3):
        out2, loss2, g2 = O.motion_guidance_f32(z, ff, fb, fo, bo, 32.0, want_loss=True)
        assert torch.equal(out2, out) and torch.equal(g2, g) and torch.equal(loss2, loss)


def test_canvas_posterior():
    O = ops()
    from oracle.torch_ref import canvas_tiles, gaussian_weights
    T, C, h, w, ts = 2, 4, 92, 120, 64
    x, noise = rnd(T, C, h, w), rnd(T, C, h, w, seed=5)
    offs = canvas_tiles(h, w, ts, 32)
    tiles = [rnd(T, C, ts, ts, seed=10 + i) for i in range(len(offs))]
    tw = gaussian_weights(ts, ts, 1)[0, 0].to(DEV)
    args = (offs, ts, 1.7, 1.3, 0.4, 0.6, 0.25)
    got, ge = O.canvas_posterior_f32(x, tiles, tw, noise, *args, want_eps=True)
    ref, re_ = E.canvas_posterior_f32(x, tiles, tw, noise, *args, want_eps=True)
    assert rel_err(ge, re_) < 1e-6 and rel_err(got, ref) < 1e-6


def test_raft_ops():
    O = ops()
    # 1x5 / 5x1 taps + sigmoid / tanh epilogues, two sources (the SepConvGRU shape)
    net, x = rnd(4, 16, 17, 128).half(), rnd(4, 16, 17, 256).half()
    for taps, kk in ((O.TAPS_1X5, (1, 5)), (O.TAPS_5X1, (5, 1))):
        w = O.pack_conv_weight(rnd(256, 384, *kk, scale=(5 * 384) ** -0.5))
        b = rnd(256)
        for act in (O.ACT_SIGMOID, O.ACT_TANH):
            kw = dict(taps=taps, a2=x, bias=b, act=act)
            assert rel_err(O.conv_gemm(net, w, **kw), E.conv_gemm(net, w, **kw)) < 4e-3
    # stems, subsample, instance norm
    img = rnd(3, 3, 40, 56)
    w7, b7 = rnd(64, 3, 7, 7, scale=0.08), rnd(64)
    assert rel_err(O.conv_direct(img, w7, b7, stride=2, pad=3, relu=True), E.conv_direct(img, w7, b7, stride=2, pad=3, relu=True)) < 2e-3
    xh = rnd(3, 20, 28, 128).half()
    assert torch.equal(O.subsample2(xh), E.subsample2(xh))
    for relu in (False, True):
        assert rel_err(O.instance_norm(xh, relu=relu), E.instance_norm(xh, relu=relu)) < 3e-3
    assert rel_err(O.axpby(xh, xh.flip(0), 1.0, 1.0, relu=True), E.axpby(xh, xh.flip(0), 1.0, 1.0, relu=True)) < 2e-3
    # correlation pyramid + lookup
    B, h, w = 2, 16, 17
    corr = rnd(B * h * w, h, w)
    levels = [corr]
    for _ in range(3):
        levels.append(O.avgpool2_f32(levels[-1]))
        assert rel_err(levels[-1], E.avgpool2_f32(levels[-2])) < 1e-6
    ys, xs = torch.meshgrid(torch.arange(h, device=DEV), torch.arange(w, device=DEV), indexing="ij")
    coords = torch.stack([xs, ys], 0).float()[None].repeat(B, 1, 1, 1) + rnd(B, 2, h, w) * 2
    got = O.corr_lookup(levels, coords, torch.zeros(B, h, w, 384, device=DEV, dtype=torch.float16))
    ref = E.corr_lookup([l.cpu() for l in levels], coords.cpu(), torch.zeros(B, h, w, 384, dtype=torch.float16))
    assert rel_err(got.cpu(), ref) < 2e-3 and got[..., 324:].abs().max() == 0
    # GRU gating, channel insert, convex upsampling
    zr, q, nt = rnd(B, h, w, 256).sigmoid().half(), rnd(B, h, w, 128).tanh().half(), rnd(B, h, w, 128).half()
    assert rel_err(O.gru_rh(zr, nt), E.gru_rh(zr, nt)) < 2e-3
    n1, n2 = nt.clone(), nt.clone()
    assert rel_err(O.gru_update(zr, q, n1), E.gru_update(zr, q, n2)) < 2e-3
    fl = rnd(B, 2, h, w)
    buf1, buf2 = torch.zeros(B, h, w, 256, device=DEV, dtype=torch.float16), torch.zeros(B, h, w, 256, device=DEV, dtype=torch.float16)
    assert torch.equal(O.set_channels(fl, buf1, 254), E.set_channels(fl, buf2, 254))
    mask = rnd(B, h, w, 576).half()
    assert rel_err(O.convex_upsample8(mask, fl), E.convex_upsample8(mask, fl)) < 1e-4


@pytest.mark.parametrize("n,h,w,scale", [(3, 128, 128, 4.0), (2, 180, 320, 4.0), (2, 96, 150, 512 / 96), (1, 270, 480, 4.0)])
def test_io_edges(n, h, w, scale):
    """GPU-side I/O edges: read_image + bicubic (+ clamp + reflect pad) and the uint8 quantisation, against
    the torch / numpy operations the reference script performs (script :124-130, :349-357, :376, :383-387, :529-541)"""
    O = ops()
    g = torch.Generator().manual_seed(n * h + w)
    u8 = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8).to(DEV)
    oh, ow = int(h * scale), int(w * scale)
    ref = E.frames_u8_to_f32_bicubic(u8, oh, ow)
    got = O.frames_u8_to_f32_bicubic(u8, oh, ow)
    assert got.shape == ref.shape and (got - ref).abs().max() < 2e-5
    ph, pw = ((oh // 32) + 1) * 32 - oh, ((ow // 32) + 1) * 32 - ow
    ref = E.frames_u8_to_f32_bicubic(u8, oh, ow, ph, pw, clamp=True)
    got = O.frames_u8_to_f32_bicubic(u8, oh, ow, ph, pw, clamp=True)
    assert got.shape == ref.shape and (got - ref).abs().max() < 2e-5
    sr = torch.rand(n, 3, oh + ph, ow + pw, generator=g).to(DEV)
    sr[0, :, :4, :4] = torch.tensor([0.0, 1.0, 254.999 / 255, 0.5]).to(DEV)
    assert torch.equal(O.frames_f32_to_u8_hwc(sr, oh, ow), E.frames_f32_to_u8_hwc(sr, oh, ow))
    assert torch.equal(O.frames_f32_to_u8_hwc(sr), E.frames_f32_to_u8_hwc(sr))


def test_stats_pool_after_graph_replay():
    """Regression (r02): the pooled GroupNorm accumulators are zeroed lazily by Python-side dirty flags; a CUDA-graph REPLAY
    writes its slots without touching the flags.  An eager forward that was the first user of a key after (i) an eager
    reset had cleared the flags and (ii) a replayed graph of another forward had written the key's slots accumulated onto
    stale sums (first unit of a clip wrong after a VAE-tiled clip had been processed).  Graph owners now mark the pool dirty
    after every replay."""
    O = ops()
    x, y = (rnd(4, 1024, 64) + 0.3).half(), (rnd(2, 1024, 64, seed=5) - 0.2).half()
    g, b = rnd(64), rnd(64, seed=2)

    def forward(t):                       # what a model forward does: one reset, then its normalisations
        O.stats_pool_reset()
        return O.group_norm(t, g, b, 1e-5, True)
    ref = E.group_norm(x, g, b, 1e-5, True)
    assert rel_err(forward(x), ref) < 3e-3
    f = O.GraphedFn(forward)
    assert rel_err(f(x), ref) < 3e-3      # warm-up + capture + replay
    forward(y)                            # an eager forward on another key: its reset() clears every flag
    assert rel_err(f(x), ref) < 3e-3      # replay: the device slots of x's key are written, Python flags are not
    assert rel_err(forward(x), ref) < 3e-3, "eager forward accumulated onto the sums left by a graph replay"
    assert rel_err(forward(x), ref) < 3e-3
