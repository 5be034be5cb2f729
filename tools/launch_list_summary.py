"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active...,sm__inst_executed_pipe_xu... --csv`) of one eager tile-step: per kernel family the launch
count, total duration, share of the step, DRAM bytes and duration-weighted tensor / XU pipe activity.  Also (re)writes
profiles/conv_gemm_ncu_traffic.json, which bench.py reports as `roofline.traffic`.

    python tools/launch_list_summary.py gpurun_out/r02_ncu_launches_tile_step_T10.csv profiles/r02_ncu_launch_shares.txt
"""
import collections, csv, json, os, re, sys

src, dst = sys.argv[1], sys.argv[2]
rows = []
with open(src, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
iid, iname, imet, ival = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
per = collections.OrderedDict()
for r in rd:
    d = per.setdefault(r[iid], {"name": r[iname]})
    try:
        d[r[imet]] = float(r[ival].replace(",", ""))
    except ValueError:
        pass

def family(n):
    n = re.sub(r"^void ", "", n)
    m = re.match(r"(?:mgld::)?(\w+)", n)
    base = m.group(1) if m else n
    return ("mgld::" if "mgld::" in n else "") + base

agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(family(d["name"]), dict(n=0, ns=0.0, rd=0.0, wr=0.0, tp=0.0, xu=0.0))
    ns = d.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1; a["ns"] += ns
    a["rd"] += d.get("dram__bytes_read.sum", 0.0); a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    a["tp"] += ns * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    a["xu"] += ns * d.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 0.0)
tot = sum(a["ns"] for a in agg.values())
out = [f"# {src}: {len(per)} launches, {tot / 1e6:.3f} ms serialised (cold-cache, under ncu: compare SHARES, not absolutes)",
       f"{'kernel':44s} {'n':>5s} {'ms':>8s} {'share':>7s} {'dram MB':>9s} {'tensor%':>8s} {'xu%':>6s}"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    out.append(f"{k[:44]:44s} {a['n']:5d} {a['ns'] / 1e6:8.3f} {100 * a['ns'] / tot:6.1f}% {(a['rd'] + a['wr']) / 1e6:9.1f} "
               f"{a['tp'] / max(a['ns'], 1):8.1f} {a['xu'] / max(a['ns'], 1):6.1f}")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:25]))
cg = agg.get("mgld::conv_gemm_kernel")
if cg:
    j = {"source": os.path.basename(src), "launches": cg["n"], "dram_bytes_per_launch": (cg["rd"] + cg["wr"]) / cg["n"],
         "dram_bytes_total": cg["rd"] + cg["wr"], "tensor_pipe_pct_duration_weighted": cg["tp"] / cg["ns"]}
    json.dump(j, open(os.path.join(os.path.dirname(dst), "conv_gemm_ncu_traffic.json"), "w"), indent=1)
