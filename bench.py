#!/usr/bin/env python
"""bench.py — HR frames/s at 50 DDPM steps on a synthetic 8-frame 512x512 clip (BASELINE.json config[1]).

One "step" = one pass of the hot path over one clip per GPU: bicubic x4 -> VAE-encode LR -> q_sample ->
50 x [struct-cond encoder + SD-2.1 UNet tile-step + posterior + motion guidance] -> video-VAE encode taps + temporal
decode -> AdaIN colour fix, for the 2 five-frame segments of an 8-frame clip (the last segment is padded by repeating
the last frame, script :345-346; only the 8 real frames are counted).  Weights are random-init at the reference's
architecture (no checkpoints on the box), the text context is a random (1,77,1024) tensor, flows are synthetic smooth
fields only with --flow synthetic (default: RAFT_SR runs inside the timed path like in the reference), data = synthetic.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU): weak scaling, every rank runs its own clip, and the output clips are
stitched with ONE NCCL all-gather inside the timed region.  `--impl reference` times the oracle (oracle/torch_ref.py,
the CPU restatement of the reference's PyTorch path) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.nn.functional as F

METRIC = "hr_frames_per_sec_512x512_50_ddpm_steps"
UNIT = "frames/s"
N_FRAMES_CLIP = 8
# algorithmic work (FLOP = 2*MAC, FlopCounterMode on the reference modules, SURVEY.md §8d / BASELINE.md §2)
FLOP_TILE_STEP = 4.390e12 + 0.447e12
FLOP_VAE_ENC, FLOP_VAE_DEC = 5.583e12, 20.695e12


def fast_state_dict(shapes, seed):
    """random-init weights of the reference's architecture: N(0, 1/fan_in) so activations stay O(1) through ~60 layers;
    zero-init modules of the reference (zero_module) are randomised too (else their outputs hide work)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, s in shapes.items():
        s = tuple(s)
        if k.endswith("temporal_alpha"):
            sd[k] = torch.full(s, 0.5)
        elif k.endswith("running_var"):
            sd[k] = torch.ones(s)
        elif k.endswith("running_mean") or k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros(s, dtype=torch.long if k.endswith("tracked") else torch.float32)
        elif len(s) == 1 and k.endswith(".weight"):
            sd[k] = torch.ones(s)
        elif k.endswith(".bias"):
            sd[k] = torch.zeros(s)
        else:
            fan = 1
            for d in s[1:]:
                fan *= d
            sd[k] = torch.randn(s, generator=g) * fan ** -0.5
    return sd


def load_cfg():
    from mgld_vsr_b200.config import load_config
    return load_config(os.path.join(ROOT, "configs", "mgldvsr_sd21_shapes.yaml"))


def synthetic_clip(seed, n=N_FRAMES_CLIP, h=128, w=128):
    g = torch.Generator().manual_seed(seed)
    # smooth-ish content: low-frequency field + noise, in [-1, 1]
    base = F.interpolate(torch.rand(n, 3, 16, 16, generator=g), size=(h, w), mode="bicubic", align_corners=False)
    return (base + 0.05 * torch.randn(n, 3, h, w, generator=g)).clamp(0, 1) * 2 - 1


def synthetic_flows(seed, n_seg, T, h, w):
    """smooth +-1.5 latent-pixel flows (fwd ~ -bwd + noise) so the occlusion masks are mixed"""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_seg):
        ff = 1.5 * F.interpolate(torch.randn(T - 1, 2, 8, 8, generator=g), size=(h, w), mode="bicubic")
        fb = -ff + 0.3 * F.interpolate(torch.randn(T - 1, 2, 8, 8, generator=g), size=(h, w), mode="bicubic")
        out.append((ff, fb))
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in o.split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples), "reasons": reasons}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(cfg, steps=1, warmup=0, seed=0, verbose=False):
    """Bounded CPU sample (~10-30 s): the oracle's DDPM tile-step (struct encoder + UNet + stitch + posterior) on ONE frame
    of the 5-frame segment (64x64 latent), one VAE-encoder pass and one temporal-decoder pass on one 512^2 frame.  All
    three are linear in the number of frames (batch dimension; the temporal layers are <0.3 % of the FLOPs), so the
    frames/s of the full workload is extrapolated as 2 segments x 5 frames x (50 tile-steps + 2 encoders + decoder)."""
    from oracle import torch_ref as R
    from mgld_vsr_b200.autoencoder import VideoAutoencoderKLResi
    from mgld_vsr_b200.unet import InflatedEncoderUNetModelWT, InflatedUNetModelDualcondV2
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    threads = min(cores, 32)          # torch's CPU conv/GEMM stop scaling (and regress) beyond ~32 threads at these sizes
    torch.set_num_threads(threads)
    mp = cfg.model.params
    T = mp.num_frames
    ucfg, scfg, dd = dict(mp.unet_config.params), dict(mp.structcond_stage_config.params), dict(cfg.video_vae.params.ddconfig)
    ucfg["num_frames"] = scfg["num_frames"] = dd["num_frames"] = 1
    sd_u = fast_state_dict(InflatedUNetModelDualcondV2(**ucfg).expected_shapes(), seed)
    sd_s = fast_state_dict(InflatedEncoderUNetModelWT(**scfg).expected_shapes(), seed + 1)
    sd_v = fast_state_dict(VideoAutoencoderKLResi(ddconfig=dd, embed_dim=4).expected_shapes(), seed + 2)
    _, resp, use = R.respaced_schedule(ddpm_steps=50)
    model = R.RefModel({**{"model.diffusion_model." + k: v for k, v in sd_u.items()},
                        **{"structcond_stage_model." + k: v for k, v in sd_s.items()}}, ucfg, scfg, resp, use, 1)
    g = torch.Generator().manual_seed(seed)
    x, lat = torch.randn(1, 4, 64, 64, generator=g), torch.randn(1, 4, 64, 64, generator=g)
    ctx, noise = torch.randn(1, 77, 1024, generator=g), torch.randn(1, 4, 64, 64, generator=g)
    tw = R.gaussian_weights(64, 64, 1)

    def tile_step():
        with torch.no_grad():
            return model.p_sample_canvas(x, ctx, lat, 25, noise, None, None, -10.0, 64, 32, tw)[0]

    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        tile_step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if verbose:
            print(f"[cpu] 1-frame tile-step {dt:.2f}s", file=sys.stderr, flush=True)
    t_step = T * sum(times) / len(times)
    img = torch.rand(1, 3, 512, 512, generator=g) * 2 - 1
    with torch.no_grad():
        t0 = time.perf_counter()
        mom, fea = R.video_vae_encode(sd_v, dd, img)
        t_enc = T * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        R.video_vae_decode(sd_v, dd, torch.randn(1, 4, 64, 64, generator=g), fea, 1.0)
        t_dec = T * (time.perf_counter() - t0)
    t_clip = 2 * (50 * t_step + 2 * t_enc + t_dec)
    return {"value": N_FRAMES_CLIP / t_clip, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"oracle/torch_ref.py fp32 on {threads} threads ({cores} visible): {len(times)} one-frame DDPM tile-step(s) "
                       f"(struct-enc + UNet + stitch + posterior, 64x64 latent), one one-frame VAE encoder pass and one one-frame "
                       f"temporal decoder pass at 512^2; x5 frames -> tile-step {t_step:.2f} s, encoder {t_enc:.2f} s, decoder "
                       f"{t_dec:.2f} s per 5-frame segment; extrapolated to 2 segments x (50 steps + 2 enc + dec)"),
            "tile_step_s": t_step, "vae_enc_s": t_enc, "vae_dec_s": t_dec}, t_step


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, t_step = cpu_reference_sample(cfg, steps=args.steps, warmup=args.warmup, verbose=True)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(args, 1), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"{N_FRAMES_CLIP}-frame 512x512 synthetic clip per GPU (2 segments of 5 frames, last frame padded), "
                        f"ddpm_steps={args.ddpm_steps}, SD-2.1 UNet shape (935M) + struct-cond encoder + temporal VAE, 1 UNet tile/step "
                        f"per segment, the 2 independent segments sampled in lock-step as one (b t) = 10-frame UNet batch, "
                        f"RAFT flow + occlusion masks + motion guidance on (flow={args.flow})",
            "frames_per_gpu": N_FRAMES_CLIP, "global_frames": N_FRAMES_CLIP * world, "ddpm_steps": args.ddpm_steps,
            "parallelism": f"clip-per-GPU x{world} + 1 NCCL all-gather" if world > 1 else "single GPU",
            "l2": "no flush needed: 2.3 GB of fp16 weights are re-streamed every DDPM step (>> 126 MB L2)"}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mgld", choices=["mgld", "reference"])
    ap.add_argument("--ddpm-steps", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flow", default="raft", choices=["raft", "synthetic"],
                    help="raft: RAFT_SR flow estimation inside the timed path (the reference's behaviour); synthetic: given flows")
    args = ap.parse_args()
    cfg = load_cfg()
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__
    if not os.path.exists(os.path.join(ROOT, "mgld-vsr_b200", "libmgld.so")):
        __graft_entry__.build()
    from mgld_vsr_b200 import ops
    from mgld_vsr_b200.config import instantiate_from_config
    from mgld_vsr_b200.pipeline import VSRPipeline

    # ---- build the models (random-init, reference architecture) -------------------------------------------------------
    model = instantiate_from_config(cfg.model, device=str(dev))
    vq = instantiate_from_config(cfg.video_vae)
    sd = {}
    for pre, mod, seed in (("model.diffusion_model.", model.model.diffusion_model, 0),
                           ("structcond_stage_model.", model.structcond_stage_model, 1),
                           ("first_stage_model.", model.first_stage_model, 3), ("flownet_model.", model.flownet_model, 4)):
        sd.update({pre + k: v for k, v in fast_state_dict(mod.expected_shapes(), seed).items()})
    for k in sd:   # small random flow head: ten random-init GRU iterations stay bounded and the occlusion masks stay mixed
        if k.startswith("flownet_model.update_block.flow_head.conv2"):
            sd[k] = sd[k] * 0.02
    model.load_state_dict(sd, strict=False)
    del sd
    vq.load_state_dict(fast_state_dict(vq.expected_shapes(), 2), device=str(dev))
    g = torch.Generator().manual_seed(1234)
    context = torch.randn(1, 77, 1024, generator=g).to(dev)
    model.cond_stage_model.set_embedding(context)
    pipe = VSRPipeline(model, vq, ddpm_steps=args.ddpm_steps, seed=42)
    T = cfg.model.params.num_frames
    n_seg = (N_FRAMES_CLIP + T - 1) // T
    clip_host = synthetic_clip(42 + rank).pin_memory()
    flows = None if args.flow == "raft" else [(a.to(dev), b.to(dev)) for a, b in synthetic_flows(7 + rank, n_seg, T, 64, 64)]
    out_host = torch.empty(N_FRAMES_CLIP, 3, 512, 512).pin_memory()
    gather_buf = torch.empty(world * N_FRAMES_CLIP, 3, 512, 512, dtype=torch.uint8, device=dev) if world > 1 else None

    def run_clip(clip_dev):
        sr = pipe(clip_dev, context=context, flows_override=flows)
        if world > 1:   # stitch the global clip: one all-gather of the finished uint8 frames
            dist.all_gather_into_tensor(gather_buf, (sr * 255.0).round().to(torch.uint8))
        return sr

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    clip_dev = clip_host.to(dev, non_blocking=True)
    for _ in range(args.warmup):
        sr = run_clip(clip_dev)
    assert torch.isfinite(sr).all(), "non-finite output"

    # ---- device-resident throughput ---------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = ops.LAUNCHES[0]
    ms = timed(lambda: run_clip(clip_dev), args.steps)
    launches = ops.LAUNCHES[0] - n0
    # ---- end to end: pinned host LR clip -> H2D -> pipeline -> D2H of the SR frames -------------------------------------
    def e2e_step():
        d = clip_host.to(dev, non_blocking=True)
        out_host.copy_(run_clip(d), non_blocking=True)
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    value = world * N_FRAMES_CLIP * args.steps / (ms / 1e3)
    e2e_value = world * N_FRAMES_CLIP * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel (conv_gemm): per-launch CUDA events over one eager tile-step -----------------
    roofline = measure_conv_gemm_roofline(model, ops, dev, context, T * min(pipe.clips_per_batch, n_seg))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 (fp32 accumulate; fp32 norms/softmax/schedule/guidance)", "data": "synthetic",
            "config": workload_config(args, world), "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": clip_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roofline}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_reference_sample(cfg, steps=1, warmup=0)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_conv_gemm_roofline(model, ops, dev, context, T):
    """Every mgld_conv_gemm launch of one eager struct-encoder + UNet tile-step (T frames = the `(b t)` batch the timed
    workload runs per DDPM step) is bracketed by CUDA events on the
    launching stream; achieved = sum(algorithmic FLOPs) / sum(durations).  Peak: MEASURED_PEAKS.json (sustained: the
    kernel runs inside a long step), else the B200_PROFILING.md fallback."""
    peak, src = 1590.0, "fallback 1.59 PFLOP/s (B200_PROFILING.md)"
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        try:
            j = json.load(open(pk))
            peak, src = float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "MEASURED_PEAKS.json bf16_tflops_sustained"
        except Exception:
            pass
    rec, orig = [], ops.conv_gemm

    def wrapped(a, w, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(a, w, **kw)
        e1.record()
        m = a.numel() // a.shape[-1]
        rec.append((e0, e1, 2.0 * m * w.shape[0] * w.shape[1]))
        return out

    x, lat = torch.randn(T, 4, 64, 64, device=dev), torch.randn(T, 4, 64, 64, device=dev)
    t = torch.tensor([500], device=dev)
    ops.conv_gemm = wrapped
    try:
        for _ in range(2):
            rec.clear()
            model.model.diffusion_model(x, t, context=context, struct_cond=model.structcond_stage_model(lat, t))
        torch.cuda.synchronize()
    finally:
        ops.conv_gemm = orig
    tot_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in rec)
    tot_fl = sum(f for _, _, f in rec)
    achieved = tot_fl / (tot_ms * 1e-3) / 1e12
    traffic = None
    tj = os.path.join(ROOT, "profiles", "conv_gemm_ncu_traffic.json")
    if os.path.exists(tj):
        try:
            traffic = json.load(open(tj)).get("dram_bytes_per_launch")
        except Exception:
            pass
    return {"kernel": "mgld::conv_gemm_kernel (tcgen05 implicit-GEMM conv/linear)", "bound": "tensor", "achieved": achieved,
            "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "peak_source": src,
            "launches_timed": len(rec), "flop_per_tile_step_in_kernel": tot_fl, "frames_per_tile_step": T,
            "note": "achieved = sum of algorithmic FLOPs of the %d conv_gemm launches of one struct-enc+UNet tile-step / "
                    "sum of their CUDA-event durations (eager launches: includes ~2 us of event/launch gap each)" % len(rec)}


if __name__ == "__main__":
    main()
