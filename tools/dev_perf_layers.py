"""GPU dev perf: the conv_gemm layers of one T=10 tile-step that sit furthest below the tensor roofline (list from
tools/trace_layer_shapes.py), each timed alone with CUDA events over rotating buffer sets (so that nothing stays in L2
that would not in the network), plus a checksum of the output to compare kernel variants across processes.

    MGLD_CONV_PAIR_EPI_BN=256 MGLD_CONV_RASTER=1 python tools/dev_perf_layers.py [T]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgld_vsr_b200 import ops as O
dev = "cuda"
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
g = torch.Generator(device=dev); g.manual_seed(0)
rnd = lambda *s, scale=1.0: torch.randn(*s, device=dev, generator=g) * scale


def bench(fns, n=24, reps=3):
    """n launches captured in one CUDA graph (eager launches of 10-20 us kernels would time the host), replayed reps times"""
    for f in fns: f()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(n): fns[i % len(fns)]()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3


LAYERS = [  # (kind, H, C_in, N_out, taps)
    ("geglu", 64, 320, 1280, 1), ("geglu", 32, 640, 2560, 1), ("geglu", 16, 1280, 5120, 1), ("geglu", 8, 1280, 5120, 1),
    ("spade", 64, 128, 320, 9), ("spade", 32, 128, 640, 9), ("spade", 16, 128, 1280, 9), ("spade", 8, 128, 1280, 9),
    ("res", 64, 320, 320, 1), ("res", 32, 640, 640, 1), ("res", 16, 1280, 1280, 1),
    ("lin", 64, 320, 320, 1), ("lin", 64, 320, 960, 1), ("lin", 32, 640, 1920, 1),
    ("res", 64, 1280, 320, 1), ("res", 32, 2560, 640, 1), ("res", 16, 5120, 1280, 1),
    ("lin", 64, 320, 320, 9), ("lin", 64, 960, 320, 9), ("lin", 32, 640, 640, 9), ("lin", 16, 1280, 1280, 9),
    ("lin", 16, 2560, 1280, 9), ("lin", 8, 1280, 1280, 9), ("lin", 8, 2560, 1280, 9), ("res", 8, 1280, 1280, 1), ("res", 8, 5120, 1280, 1),
]
tot = 0.0
for kind, H, Ci, No, taps in LAYERS:
    M = T * H * H
    out_bytes = M * No * 2; in_bytes = M * Ci * 2
    nset = max(2, min(8, int(300e6 // (out_bytes + in_bytes)) + 1))
    sets = []
    for k in range(nset):
        a = rnd(T, H, H, Ci).half()
        if kind == "geglu":
            w = O.interleave_pair(rnd(No, Ci, scale=Ci ** -0.5).half(), rnd(No, Ci, scale=Ci ** -0.5).half())
            kw = dict(bias=rnd(2 * No), epilogue=O.EPI_GEGLU)
        elif kind == "spade":
            w = O.interleave_pair(O.pack_conv_weight(rnd(No, Ci, 3, 3, scale=0.03)), O.pack_conv_weight(rnd(No, Ci, 3, 3, scale=0.03)))
            h = rnd(T, H, H, No).half()
            st = O.gn_finalize(O.gn_stats(h), H * H, No, 1e-5)
            kw = dict(taps=9, bias=rnd(2 * No, scale=0.1), epilogue=O.EPI_SPADE, h=h, gn_stats=st, gn_weight=rnd(No), gn_bias=rnd(No),
                      groups=32, res=rnd(T, H, H, No).half(), beta=1.0)
        else:
            w = rnd(No, taps * Ci, scale=(taps * Ci) ** -0.5).half()
            kw = dict(taps=taps, bias=rnd(No))
            if kind == "res":
                kw.update(res=rnd(T, H, H, No).half(), beta=1.0)
        if k and kind != "spade":
            w = sets[0][1]; kw["bias"] = sets[0][2]["bias"]      # weights are shared by all sets (they do live in L2)
        out = torch.empty(T, H, H, No, device=dev, dtype=torch.float16)
        sets.append((a, w, kw, out))
    fns = [(lambda s=s: O.conv_gemm(s[0], s[1], out=s[3], **s[2])) for s in sets]
    us = bench(fns)
    Nw = sets[0][1].shape[0]
    fl = 2.0 * M * Nw * taps * Ci
    o = sets[0][3]
    tot += us
    print(f"{kind:5s} {H:2d}x{H:<2d} {Ci:5d}->{No:5d} taps{taps}: {us:7.1f} us {fl / us / 1e6:7.0f} TFLOP/s  sets {nset}  "
          f"chk {o.float().abs().sum().item():.6e} {int(o.view(torch.int16).long().sum().item())}", flush=True)
print(f"sum {tot:.1f} us")
