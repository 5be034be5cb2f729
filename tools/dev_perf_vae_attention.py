"""GPU dev perf: the VAEs' middle attention block (single head, d = 512; autoencoder._AttnBlock) at a 512x512 frame batch
(N = 4096, T = 5) and at a 960x960 VAE tile (N = 14400, T = 2): the fused path (one q|k|v GEMM + the split-D flash kernel,
csrc/attention_hd512.cu, with 4 / 2 slabs per TMA operation) against the panelled GEMM -> softmax -> GEMM path, the flash
kernel alone, and the agreement of the two block outputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgld_vsr_b200 import ops
from mgld_vsr_b200.autoencoder import _AttnBlock
from mgld_vsr_b200.unet import _Packed

C = 512
g = torch.Generator().manual_seed(0)
sd = {}
for n in ("q", "k", "v", "proj_out"):
    sd[f"a.{n}.weight"] = torch.randn(C, C, 1, 1, generator=g) * C ** -0.5
    sd[f"a.{n}.bias"] = torch.randn(C, generator=g) * 0.1
sd["a.norm.weight"], sd["a.norm.bias"] = torch.ones(C), torch.zeros(C)
blk = _AttnBlock(_Packed(sd, torch.device("cuda")), "a")


def bench(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run_block(x):
    ops.stats_pool_reset()
    return blk(ops, x)


for (T, HW) in [(5, 64), (2, 120)]:
    N = HW * HW
    x = torch.randn(T, HW, HW, C, generator=g).half().cuda()
    qkv = torch.randn(T * N, 3 * C, generator=g).half().cuda()
    out = torch.empty(T * N, C, device="cuda", dtype=torch.float16)
    flop = 4.0 * T * N * N * C
    res = {}
    for grp in ("4", "2"):
        os.environ["MGLD_HD512_GROUP"] = grp
        ms = bench(lambda: ops.attention(qkv, qkv, qkv, batch=T, heads=1, head_dim=C, nq=N, nkv=N, scale=C ** -0.5,
                                         q_col0=0, k_col0=C, v_col0=2 * C, out=out))
        print(f"T{T} N{N} flash kernel, {grp} slab(s) per TMA op: {ms * 1e3:.0f} us = {ms / T * 1e3:.0f} us/frame, "
              f"{flop / ms / 1e9:.0f} TFLOP/s algorithmic", flush=True)
    os.environ["MGLD_HD512_GROUP"] = "4"
    _AttnBlock.FUSED_HEAD_DIMS = (64, 128, 512)
    res["fused"] = (bench(lambda: run_block(x)), run_block(x).float())
    _AttnBlock.FUSED_HEAD_DIMS = ()
    res["panelled"] = (bench(lambda: run_block(x)), run_block(x).float())
    d = (res["fused"][1] - res["panelled"][1]).abs().max() / res["panelled"][1].abs().max()
    print(f"T{T} N{N} whole block: fused {res['fused'][0] * 1e3:.0f} us, panelled {res['panelled'][0] * 1e3:.0f} us "
          f"({res['panelled'][0] / res['fused'][0]:.2f}x); outputs differ by {d:.2e} of range", flush=True)
