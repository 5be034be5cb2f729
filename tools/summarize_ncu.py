"""Summarise an `ncu --set full --import-source on` report into a small text file for profiles/:
per-kernel headline metrics (duration, tensor / XU / FMA / ALU pipe utilisation, issue-slot utilisation, DRAM bytes) and, for
the first kernel, the warp-stall sampling by reason and the most-stalled SASS lines.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/out.txt
"""
import collections, csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("duration", "gpu__time_duration.sum"), ("sm_clock", "sm__cycles_elapsed.avg.per_second"),
        ("tensor_pipe_active_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("fma_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        ("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        ("issue_active_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        ("dram_read", "dram__bytes_read.sum"), ("dram_write", "dram__bytes_write.sum"),
        ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("lts_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("registers", "launch__registers_per_thread"), ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active")]
lines = [f"# {rep}: ncu --set full --clock-control none (values under the profiler; durations are not bench values)", ""]
for r in data:
    lines.append(f"kernel {r[hdr.index('Kernel Name')].split('(')[0]}  grid {r[hdr.index('Grid Size')]}  block {r[hdr.index('Block Size')]}")
    for name, col in want:
        if col in hdr:
            i = hdr.index(col)
            lines.append(f"    {name:26s} {r[i]} {units[i]}")
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
# the source page repeats a 2-line header per kernel; take the first kernel
if len(srows) > 2:
    shdr = srows[1]
    body = []
    for r in srows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(shdr):
            body.append(r)
    isamp, isrc, iex = shdr.index("# Samples"), shdr.index("Source"), shdr.index("Instructions Executed")
    stall = [(i, h) for i, h in enumerate(shdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp] or 0) for r in body)
    agg = collections.Counter()
    for r in body:
        for i, h in stall:
            if r[i] not in ("", "0"):
                agg[h] += int(r[i])
    lines.append(f"warp-state samples of the first kernel: {tot}")
    lines.append("    by reason: " + ", ".join(f"{h[6:]} {n}" for h, n in agg.most_common(10)))
    ops = collections.Counter()
    for r in body:
        t = r[isrc].split()
        if t:
            ops[t[1] if t[0].startswith("@") and len(t) > 1 else t[0]] += int(r[isamp] or 0)
    lines.append("    by opcode: " + ", ".join(f"{o} {n}" for o, n in ops.most_common(12)))
    lines.append("    most-stalled SASS lines (samples, executed, instruction, top reasons):")
    for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:25]:
        st = sorted([(int(r[i]), h[6:]) for i, h in stall if r[i] not in ("", "0")], reverse=True)[:2]
        lines.append(f"      {int(r[isamp] or 0):5d} {r[iex]:>9s}  {r[isrc][:70]:70s} {st}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
