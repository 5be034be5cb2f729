"""GPU dev: per-call CUDA-event timing of every op of one eager struct-encoder + UNet tile-step, grouped by (op, shape).
MGLD_T = frames in the batch (5 = one clip, 10 = two clips)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench import fast_state_dict, load_cfg
from mgld_vsr_b200 import ops
from mgld_vsr_b200.config import instantiate_from_config
cfg = load_cfg(); dev = "cuda"
mp = cfg.model.params
unet = instantiate_from_config(mp.unet_config); se = instantiate_from_config(mp.structcond_stage_config)
unet.load_state_dict(fast_state_dict(unet.expected_shapes(), 0)); se.load_state_dict(fast_state_dict(se.expected_shapes(), 1))
T = int(os.environ.get("MGLD_T", "10"))
x = torch.randn(T, 4, 64, 64, device=dev); lat = torch.randn(T, 4, 64, 64, device=dev)
ctx = torch.randn(1, 77, 1024, device=dev); t = torch.tensor([500], device=dev)
rec = []
def wrap(name, keyfn):
    orig = getattr(ops, name)
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = orig(*a, **k); e1.record()
        rec.append((name, keyfn(*a, **k), e0, e1))
        return out
    setattr(ops, name, f)
def conv_key(a, w, **k):
    M = a.numel() // a.shape[-1]
    C = a.shape[-1] + (k["a2"].shape[-1] if k.get("a2") is not None else 0)
    taps = k.get("taps", 1)
    return (M, w.shape[0], C * (5 if taps == 6 else taps), taps, k.get("epilogue", 0), k.get("act", 0), k.get("res") is not None)
wrap("conv_gemm", conv_key)
wrap("attention", lambda q, k_, v, **k: (k["batch"], k["heads"], k["head_dim"], k["nq"], k["nkv"]))
wrap("group_norm", lambda x1, *a, **k: (tuple(x1.shape), k.get("x2").shape[-1] if k.get("x2") is not None else 0, k.get("want_out", True)))
wrap("layernorm", lambda x_, *a, **k: tuple(x_.shape))
wrap("conv_small_cin", lambda x_, w, b: (tuple(x_.shape), w.shape[0]))
wrap("conv3x3_small_cout", lambda x_, w, b: tuple(x_.shape))
wrap("im2col_s2", lambda x_, p: tuple(x_.shape))
wrap("upsample2x", lambda x_: tuple(x_.shape))
wrap("gemv", lambda x_, w, **k: tuple(w.shape))
for _ in range(3):
    rec.clear()
    out = unet(x, t, ctx, se(lat, t))
torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, key, e0, e1 in rec:
    a = agg.setdefault((name, key), [0, 0.0])
    a[0] += 1; a[1] += e0.elapsed_time(e1) * 1e3
tot = sum(v[1] for v in agg.values())
print(f"T={T}: {len(rec)} calls, sum of per-call times {tot/1e3:.2f} ms (eager, includes launch gaps)")
byop = collections.Counter()
for (name, key), (n, us) in agg.items(): byop[name] += us
print({k: round(v / 1e3, 2) for k, v in byop.most_common()})
print(f"{'op':14s} {'shape':58s} {'n':>3s} {'us_tot':>8s} {'us_each':>8s} {'TFLOP/s':>8s} {'%':>5s}")
for (name, key), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    tf = ""
    if name == "conv_gemm": tf = f"{2.0 * key[0] * key[1] * key[2] * n / us / 1e6:8.0f}"
    if name == "attention": tf = f"{4.0 * key[0] * key[1] * key[2] * key[3] * key[4] * n / us / 1e6:8.0f}"
    print(f"{name:14s} {str(key):58s} {n:3d} {us:8.1f} {us / n:8.1f} {tf:>8s} {100 * us / tot:5.1f}")
