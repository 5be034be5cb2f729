"""List the conv_gemm launches of one struct-encoder + UNet tile-step (SD-2.1 shapes) in launch order with their GEMM
shapes, by running the host graph on the CPU through tests/emu_ops.py with a logging wrapper; optionally join the list
with an ncu launch list of tools/ncu_target.py (same order) to see which layers lose the tensor pipe.

    python tools/trace_layer_shapes.py [T] [ncu_launches.csv] > profiles/..._by_layer.txt
"""
import collections, csv, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import emu_ops
from bench import fast_state_dict, load_cfg
from mgld_vsr_b200.config import instantiate_from_config

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
LOG = []
ops = types.ModuleType("trace_ops")
ops.__dict__.update(emu_ops.__dict__)
EPI = {0: "lin", 1: "geglu", 2: "spade"}


def conv_gemm(a, w, *, taps=1, a2=None, bias=None, epilogue=0, act=0, res=None, out_f32=False, **kw):
    M = a.numel() // a.shape[-1]
    K = w.shape[1]
    LOG.append(dict(M=M, N=w.shape[0], K=K, taps=taps, epi=EPI[epilogue], act=act, res=res is not None, a2=a2 is not None,
                    f32=bool(out_f32), hw=tuple(a.shape[1:-1])))
    return emu_ops.conv_gemm(a, w, taps=taps, a2=a2, bias=bias, epilogue=epilogue, act=act, res=res, out_f32=out_f32, **kw)


def conv3x3_small_cout(x, w_packed, bias):      # ops.conv3x3_small_cout runs conv_gemm (N=32, fp32 out) + a channel slice
    LOG.append(dict(M=x.numel() // x.shape[-1], N=32, K=9 * x.shape[-1], taps=9, epi="lin", act=0, res=False, a2=False, f32=True,
                    hw=tuple(x.shape[1:-1])))
    return emu_ops.conv3x3_small_cout(x, w_packed, bias)


ops.conv_gemm = conv_gemm
ops.conv3x3_small_cout = conv3x3_small_cout
cfg = load_cfg(); mp = cfg.model.params
for c in (mp.unet_config, mp.structcond_stage_config):
    c.params["ops"] = ops
unet = instantiate_from_config(mp.unet_config); se = instantiate_from_config(mp.structcond_stage_config)
unet.load_state_dict(fast_state_dict(unet.expected_shapes(), 0), device="cpu")
se.load_state_dict(fast_state_dict(se.expected_shapes(), 1), device="cpu")
x = torch.randn(T, 4, 64, 64); lat = torch.randn(T, 4, 64, 64); ctx = torch.randn(1, 77, 1024); t = torch.tensor([500])
with torch.no_grad():
    unet(x, t, ctx, se(lat, t))
    LOG.clear()                      # the first call also runs the cached text K/V GEMMs
    unet(x, t, ctx, se(lat, t))

ncu = None
if len(sys.argv) > 2:
    lines = [l for l in open(sys.argv[2], newline="") if l.startswith('"')]
    rd = csv.reader(lines); hdr = next(rd)
    iid, iname, imet, ival = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    per = collections.OrderedDict()
    for r in rd:
        d = per.setdefault(r[iid], {"name": r[iname]})
        try: d[r[imet]] = float(r[ival].replace(",", ""))
        except ValueError: pass
    ncu = [d for d in per.values() if "conv_gemm_kernel" in d["name"]]
    print(f"# {len(LOG)} traced conv_gemm calls, {len(ncu)} conv_gemm launches in {sys.argv[2]}")
    assert len(ncu) == len(LOG), "launch lists differ"

agg = collections.OrderedDict()
for i, l in enumerate(LOG):
    key = (l["M"], l["N"], l["K"], l["taps"], l["epi"], l["res"], l["f32"])
    a = agg.setdefault(key, dict(n=0, ns=0.0, tp=0.0, dram=0.0))
    a["n"] += 1
    if ncu:
        ns = ncu[i]["gpu__time_duration.sum"]
        a["ns"] += ns; a["tp"] += ns * ncu[i].get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        a["dram"] += ncu[i].get("dram__bytes_read.sum", 0) + ncu[i].get("dram__bytes_write.sum", 0)
print(f"{'M':>7s} {'N':>5s} {'K':>6s} taps {'epi':>5s} res f32 {'n':>3s} {'us/launch':>9s} {'total ms':>8s} {'TFLOP/s':>8s} {'tensor%':>7s} {'MB/launch':>9s}")
tot = 0.0
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    M, N, K, taps, epi, res, f32 = k
    fl = 2.0 * M * N * K
    us = a["ns"] / a["n"] / 1e3 if ncu else 0.0
    tot += a["ns"]
    print(f"{M:7d} {N:5d} {K:6d} {taps:4d} {epi:>5s} {int(res):3d} {int(f32):3d} {a['n']:3d} {us:9.1f} {a['ns'] / 1e6:8.3f} "
          f"{(fl / (us * 1e-6) / 1e12 if us else 0):8.0f} {(a['tp'] / a['ns'] if a['ns'] else 0):7.1f} {a['dram'] / a['n'] / 1e6:9.1f}")
print(f"# total {tot / 1e6:.3f} ms")
