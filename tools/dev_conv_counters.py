"""GPU dev: where do the conv_gemm roles wait?  Per-CTA cycle counters (mgld_conv_gemm_set_debug_counters)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mgld_vsr_b200 import ops, lib as L
dev = "cuda"
import os
shapes_t10 = [(10, 64, 64, 320, 320, 1, 0), (10, 64, 64, 320, 960, 1, 0), (10, 64, 64, 1280, 320, 1, 0), (10, 32, 32, 640, 640, 1, 0),
              (10, 16, 16, 1280, 1280, 1, 0), (10, 64, 64, 320, 320, 9, 0), (10, 16, 16, 1280, 1280, 9, 0), (10, 8, 8, 1280, 1280, 9, 0)]
shapes = [(5, 64, 64, 320, 320, 9, 0), (5, 64, 64, 320, 320, 1, 0), (5, 64, 64, 320, 960, 1, 0), (5, 64, 64, 320, 2560, 1, 0),
          (5, 64, 64, 320, 2560, 1, 128), (5, 64, 64, 1280, 320, 1, 0), (5, 32, 32, 640, 640, 1, 0), (5, 32, 32, 640, 5120, 1, 0),
          (5, 16, 16, 1280, 1280, 1, 0), (5, 256, 256, 256, 256, 9, 0)]
if os.environ.get("MGLD_T") == "10": shapes = shapes_t10
if os.environ.get("MGLD_KIND") == "geglu":      # FF1 layers: N = 2 x Cout interleaved, GEGLU epilogue
    shapes = [(10, 64, 64, 320, 2560, 1, 0), (10, 32, 32, 640, 5120, 1, 0), (10, 16, 16, 1280, 10240, 1, 0)]
so = L.lib()
so.mgld_conv_gemm_set_debug_counters.argtypes = [ctypes.c_void_p]
so.mgld_conv_gemm_set_debug_counters.restype = None
for pair in (sys.argv[1:] or ["0"]):
    os.environ["MGLD_CONV_PAIR"] = pair
    for (T, H, W, Ci, Co, taps, bn) in shapes:
        x = torch.randn(T, H, W, Ci, device=dev).half(); w = (torch.randn(Co, taps * Ci, device=dev) * 0.02).half()
        b = torch.randn(Co, device=dev)
        out = torch.empty(T, H, W, Co, device=dev, dtype=torch.float16)
        kw = {}
        if os.environ.get("MGLD_KIND") == "geglu":
            kw = dict(epilogue=ops.EPI_GEGLU); out = torch.empty(T, H, W, Co // 2, device=dev, dtype=torch.float16)
        for _ in range(2): ops.conv_gemm(x, w, taps=taps, bias=b, out=out, block_n=bn, **kw)
        dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
        so.mgld_conv_gemm_set_debug_counters(ctypes.c_void_p(dbg.data_ptr()))
        ops.conv_gemm(x, w, taps=taps, bias=b, out=out, block_n=bn, **kw)
        torch.cuda.synchronize()
        so.mgld_conv_gemm_set_debug_counters(ctypes.c_void_p(0))
        d = dbg.view(148, 16).double()
        act = d[:, 7] > 0
        m = d[act].mean(0)
        ld = d[::2][d[::2, 4] > 0].mean(0) if pair == "1" else m   # MMA counters live in the leader CTAs
        nk = taps * ((Ci + 63) // 64)
        print(f"pair={pair} T{T} {H}x{W} {Ci}->{Co} taps{taps} bn{bn}: tiles/CTA {m[7]:.2f} chunks/tile {nk} | "
              f"A-prod wait_empty {m[0]:.0f} of {m[1]:.0f} | MMA wait_full {ld[2]:.0f} wait_acc {ld[3]:.0f} of {ld[4]:.0f} "
              f"({ld[4] / max(ld[7], 1) / nk:.0f} cyc/chunk) | epi wait_acc_full {m[5]:.0f} of {m[6]:.0f} ({m[6]/m[7]:.0f}/tile) | "
              f"B-prod wait_empty {m[8]:.0f} of {m[9]:.0f} | epi/tile: compute {m[10]/m[7]:.0f} drain-wait {m[11]/m[7]:.0f} | "
              f"setup {m[12]:.0f} kernel {m[13]:.0f} max {d[:,13].max():.0f}", flush=True)
