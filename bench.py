#!/usr/bin/env python
"""bench.py — HR frames/s at 50 DDPM steps (BASELINE.json metric) + UNet step ms, on synthetic clips.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|3|4|5]

One "step" = one pass of the hot path over one clip: bicubic x4 -> RAFT flow + occlusion masks -> VAE-encode LR -> q_sample
-> 50 x [struct-cond encoder + SD-2.1 UNet per 64x64 latent tile, Gaussian stitch, posterior, motion guidance] -> video-VAE
encode taps + temporal decode -> colour fix -> VAE-tile averaging, for every 5-frame segment of the clip.  Weights are
random-init at the reference's architecture (no checkpoints on the box), the text context is a random (1,77,1024) tensor.

Workloads (BASELINE.json `configs`; --config selects, default 2 = the headline the metric is quoted on):
  2  8-frame 512x512 clip (LR 128x128), 2 segments, 1 UNet tile/step.  N > 1: the reference's own multi-GPU mechanism —
     whole sequences are dealt to the ranks (`seq_idx % n_gpus == select_idx`, script :338): one clip per GPU (weak scaling),
     the finished uint8 frames of all clips are stitched with ONE NCCL all-gather inside the timed region.
  3  32-frame 512x512 clip, 7 segments, RAFT + motion guidance.
  4  16-frame 720p clip (LR 180x320 -> 736x1312 padded): 4 segments x 2 VAE tiles of 736x960, 6 UNet tiles each.
  5  64-frame 1080p clip with a synthetic real-world degradation (blur + noise): 13 segments x 6 VAE tiles of 960x960,
     9 UNet tiles each (54 tile evaluations per DDPM step per segment).
  Configs 3-5 at N > 1: ONE clip strong-scaled — its independent units (segment x VAE tile) are dealt to the ranks in
  contiguous blocks and the decoded tiles are exchanged with a single all-gather (pipeline.gather_units).

`--impl reference` times the reference's CPU path (oracle/torch_ref.py, the restatement of the reference's PyTorch code
pinned against it by tests/test_reference_pipeline.py) on the host cores, one bounded sample per step.
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

UNIT = "frames/s"
# algorithmic work (FLOP = 2*MAC, FlopCounterMode on the reference modules, SURVEY.md §8d / BASELINE.md §2)
FLOP_TILE_STEP = 4.390e12 + 0.447e12
FLOP_VAE_ENC, FLOP_VAE_DEC = 5.583e12, 20.695e12

CONFIGS = {
    2: dict(frames=8, lr=(128, 128), degrade=False, sharding="sequence", nominal_gpus=1,
            metric="hr_frames_per_sec_512x512_50_ddpm_steps",
            what="8-frame 512x512 synthetic clip (2 segments of 5 frames, last frame padded), 1 UNet tile/step per segment, "
                 "the 2 segments sampled in lock-step as one (b t) = 10-frame UNet batch"),
    3: dict(frames=32, lr=(128, 128), degrade=False, sharding="units", nominal_gpus=1,
            metric="hr_frames_per_sec_512x512_32frames_50_ddpm_steps",
            what="32-frame 512x512 synthetic clip (7 segments), 1 UNet tile/step per segment"),
    4: dict(frames=16, lr=(180, 320), degrade=False, sharding="units", nominal_gpus=4,
            metric="hr_frames_per_sec_720p_50_ddpm_steps",
            what="16-frame 720p clip (LR 180x320 -> HR 720x1280, padded 736x1312): 4 segments x 2 VAE tiles of 736x960 "
                 "(oldcanvas_tile), 6 UNet tiles per VAE tile"),
    5: dict(frames=64, lr=(270, 480), degrade=True, sharding="units", nominal_gpus=8,
            metric="hr_frames_per_sec_1080p_50_ddpm_steps",
            what="64-frame 1080p clip (LR 270x480 -> HR 1080x1920, padded 1088x1952) with synthetic real-world degradation "
                 "(box blur + Gaussian noise sigma 10/255): 13 segments x 6 VAE tiles of 960x960, 9 UNet tiles per VAE tile"),
}


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


def synthetic_clip(seed, n=8, h=128, w=128, degrade=False):
    """smooth-ish LR content in [-1, 1]; `degrade`: network-free stand-in for the real-world degradation of
    configs/mgldvsr/*.yaml:122-143: 2x box blur + Gaussian noise sigma = 10/255"""
    g = torch.Generator().manual_seed(seed)
    base = F.interpolate(torch.rand(n, 3, max(h // 8, 2), max(w // 8, 2), generator=g), size=(h, w), mode="bicubic",
                         align_corners=False)
    x = (base + 0.05 * torch.randn(n, 3, h, w, generator=g)).clamp(0, 1)
    if degrade:
        x = F.avg_pool2d(F.pad(x, (0, 1, 0, 1), mode="replicate"), 2, stride=1)
        x = (x + (10.0 / 255.0) * torch.randn(n, 3, h, w, generator=g)).clamp(0, 1)
    return x * 2 - 1


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
# the reference's path on the host cores (oracle = restatement pinned against the reference): bounded samples
# ---------------------------------------------------------------------------------------------------------------------
def _host_threads():
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    return cores, min(cores, 32)      # torch's CPU conv/GEMM stop scaling (and regress) beyond ~32 threads at these sizes


class ReferencePath:
    """The oracle's pieces of one 5-frame 512x512 segment, each callable on its own (CPU fp32, or a CUDA device under fp16
    autocast = the reference's deployment numerics): one full DDPM tile-step (struct encoder + UNet on all 5 frames,
    stitch, posterior, motion guidance), RAFT flows + masks, the three VAE passes."""

    def __init__(self, cfg, device, seed=0):
        from oracle import torch_ref as R
        from mgld_vsr_b200.autoencoder import AutoencoderKL, VideoAutoencoderKLResi
        from mgld_vsr_b200.raft import RAFT_SR
        from mgld_vsr_b200.unet import InflatedEncoderUNetModelWT, InflatedUNetModelDualcondV2
        self.R, self.dev = R, torch.device(device)
        mp = cfg.model.params
        self.T = T = mp.num_frames
        ucfg, scfg = dict(mp.unet_config.params), dict(mp.structcond_stage_config.params)
        self.dd, self.ddk = dict(cfg.video_vae.params.ddconfig), dict(mp.first_stage_config.params.ddconfig)
        to = lambda sd: {k: v.to(self.dev) for k, v in sd.items()}
        sd_u = fast_state_dict(InflatedUNetModelDualcondV2(**ucfg).expected_shapes(), seed)
        sd_s = fast_state_dict(InflatedEncoderUNetModelWT(**scfg).expected_shapes(), seed + 1)
        self.sd_v = to(fast_state_dict(VideoAutoencoderKLResi(ddconfig=self.dd, embed_dim=4).expected_shapes(), seed + 2))
        self.sd_k = to({"first_stage_model." + k: v for k, v in
                        fast_state_dict(AutoencoderKL(ddconfig=self.ddk, embed_dim=4).expected_shapes(False), seed + 3).items()})
        sd_r = fast_state_dict(RAFT_SR().expected_shapes(), seed + 4)
        for k in sd_r:
            if k.startswith("update_block.flow_head.conv2"):
                sd_r[k] = sd_r[k] * 0.02
        self.sd_r = to({"flownet_model." + k: v for k, v in sd_r.items()})
        _, resp, use = R.respaced_schedule(ddpm_steps=50)
        self.model = R.RefModel(to({**{"model.diffusion_model." + k: v for k, v in sd_u.items()},
                                    **{"structcond_stage_model." + k: v for k, v in sd_s.items()}}), ucfg, scfg, resp, use, T)
        g = torch.Generator().manual_seed(seed)
        r = lambda *s: torch.randn(*s, generator=g).to(self.dev)
        self.x, self.lat, self.ctx, self.noise = r(T, 4, 64, 64), r(T, 4, 64, 64), r(1, 77, 1024), r(T, 4, 64, 64)
        ff, fb = synthetic_flows(seed, 1, T, 64, 64)[0]
        self.flows = (ff[None].to(self.dev), fb[None].to(self.dev))
        occ = [R.forward_backward_consistency_check(self.flows[1][:, i], self.flows[0][:, i]) for i in range(T - 1)]
        self.masks = (torch.stack([o[0][:, None] for o in occ], 1), torch.stack([o[1][:, None] for o in occ], 1))
        self.tw = R.gaussian_weights(64, 64, 1).to(self.dev)
        self.img = (torch.rand(T, 3, 512, 512, generator=g) * 2 - 1).to(self.dev)
        self.z = r(T, 4, 64, 64)

    def _ctx(self):
        return torch.autocast("cuda", dtype=torch.float16) if self.dev.type == "cuda" else torch.autocast("cpu", enabled=False)

    def _sync(self):
        if self.dev.type == "cuda":
            torch.cuda.synchronize()

    def _timed(self, fn):
        self._sync()
        t0 = time.perf_counter()
        with torch.no_grad(), self._ctx():
            out = fn()
        self._sync()
        return time.perf_counter() - t0, out

    def tile_step(self):
        return self._timed(lambda: self.model.p_sample_canvas(self.x, self.ctx, self.lat, 25, self.noise, self.flows,
                                                              self.masks, -10.0, 64, 32, self.tw)[0])[0]

    def raft(self):
        lq = ((self.img + 1) / 2).clamp(0, 1)
        lq = F.interpolate(lq, size=(128, 128), mode="bicubic")[None]
        return self._timed(lambda: self.R.compute_flow(self.sd_r, lq))[0]

    def vae(self):
        t_kl = self._timed(lambda: self.R.autoencoder_kl_encode(self.sd_k, self.ddk, self.img))[0]
        t_enc, (_, fea) = self._timed(lambda: self.R.video_vae_encode(self.sd_v, self.dd, self.img))
        t_dec = self._timed(lambda: self.R.video_vae_decode(self.sd_v, self.dd, self.z, [f.float() for f in fea], 1.0))[0]
        return t_kl, t_enc, t_dec


def clip_seconds_from_units(n_segments, t_step, t_raft, t_kl, t_enc, t_dec, ddpm_steps=50):
    """a 512x512 clip processed the reference's way: segment after segment, each = RAFT + KL encode + S tile-steps + video
    encode + decode (script :375-535)"""
    return n_segments * (ddpm_steps * t_step + t_raft + t_kl + t_enc + t_dec)


def cpu_reference_sample(cfg, steps=1, warmup=0, seed=0, verbose=False):
    """Bounded CPU sample: `steps` FULL 5-frame DDPM tile-steps of the oracle (one per bench step), plus - once - RAFT and
    the three VAE passes on the full 5 frames.  frames/s of the config-2 workload = 8 / (2 segments x (50 tile-steps + RAFT +
    3 VAE passes)); nothing is extrapolated across frames."""
    cores, threads = _host_threads()
    torch.set_num_threads(threads)
    ref = ReferencePath(cfg, "cpu", seed)
    times = []
    for i in range(warmup + steps):
        dt = ref.tile_step()
        if i >= warmup:
            times.append(dt)
        if verbose:
            print(f"[cpu] 5-frame tile-step {dt:.2f}s", file=sys.stderr, flush=True)
    t_step = sum(times) / len(times)
    t_raft = ref.raft()
    t_kl, t_enc, t_dec = ref.vae()
    t_clip = clip_seconds_from_units(2, t_step, t_raft, t_kl, t_enc, t_dec)
    return {"value": 8 / t_clip, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"oracle/torch_ref.py (restatement of the reference's PyTorch path, pinned against it) fp32 on {threads} "
                       f"threads ({cores} visible): {len(times)} full 5-frame DDPM tile-step(s) (struct-enc + UNet + stitch + "
                       f"posterior + motion guidance, 64x64 latent) = {t_step:.2f} s each; once: RAFT fwd+bwd {t_raft:.2f} s, "
                       f"KL encode {t_kl:.2f} s, video encode {t_enc:.2f} s, video decode {t_dec:.2f} s on 5 frames at 512^2; "
                       f"clip = 2 segments x (50 tile-steps + those)"),
            "tile_step_s": t_step, "raft_s": t_raft, "vae_kl_enc_s": t_kl, "vae_enc_s": t_enc, "vae_dec_s": t_dec}, times


def gpu_reference_sample(cfg, dev, seed=0):
    """The north-star's real comparator ("reference-GPU"): the reference's PyTorch ops (cuDNN / cuBLAS / SDPA, eager, fp16
    autocast = its deployment numerics) on the SAME GPU, per-unit timings of one 5-frame segment composed into the config-2
    clip.  xformers is absent from the image: its fused attention is stood in for by torch's, which materialises the score
    matrix (reported, so the reader can discount it)."""
    ref = ReferencePath(cfg, dev, seed)
    for _ in range(2):
        ref.tile_step()
    ts = [ref.tile_step() for _ in range(5)]
    t_step = sum(ts) / len(ts)
    ref.raft()
    t_raft = ref.raft()
    ref.vae()
    t_kl, t_enc, t_dec = ref.vae()
    t_clip = clip_seconds_from_units(2, t_step, t_raft, t_kl, t_enc, t_dec)
    del ref
    torch.cuda.empty_cache()
    return {"value": 8 / t_clip, "unit": UNIT, "kind": "oracle torch ops (cuDNN/cuBLAS/softmax-matmul attention), eager, fp16 autocast, same GPU",
            "tile_step_ms": t_step * 1e3, "raft_ms": t_raft * 1e3, "vae_kl_enc_ms": t_kl * 1e3, "vae_enc_ms": t_enc * 1e3,
            "vae_dec_ms": t_dec * 1e3, "sample": "5 five-frame tile-steps + RAFT + 3 VAE passes, composed: 2 segments x (50 steps + RAFT + VAE)"}


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    base, times = cpu_reference_sample(cfg, steps=args.steps, warmup=min(args.warmup, 1), verbose=True)
    value = base["value"]
    if args.config != 2:   # other configs: per-unit times x unit counts; 512^2 units only compose config 3
        value = None if args.config != 3 else 32 / clip_seconds_from_units(7, base["tile_step_s"], base["raft_s"],
                                                                         base["vae_kl_enc_s"], base["vae_enc_s"], base["vae_dec_s"])
    line = {"impl": "reference", "metric": c["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "weak" if c["sharding"] == "sequence" else "strong", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "config": workload_config(args, 1), "cpu_baseline": base,
            "step_definition": "one bounded sample per step = one full 5-frame DDPM tile-step of the workload on the host cores; "
                               "`value` composes the measured unit times into the whole clip",
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    c = CONFIGS[args.config]
    frames = args.frames or c["frames"]
    seq = c["sharding"] == "sequence"
    return {"workload": f"BASELINE config {args.config}: {c['what']}; ddpm_steps={args.ddpm_steps}, SD-2.1 UNet shape (935M) + "
                        f"struct-cond encoder + temporal VAE, RAFT flow + occlusion masks + motion guidance on (flow={args.flow})",
            "baseline_config": args.config, "frames_per_clip": frames, "global_frames": frames * (world if seq else 1),
            "ddpm_steps": args.ddpm_steps,
            "parallelism": ("single GPU" if world == 1 else
                            (f"one clip per GPU x{world} (sequence sharding, script :338) + 1 NCCL all-gather of the uint8 frames" if seq
                             else f"one clip, units (segment x VAE tile) in contiguous blocks over {world} ranks + 1 NCCL all-gather of the decoded tiles")),
            "l2": "no flush needed: 2.3 GB of fp16 weights are re-streamed every DDPM step (>> 126 MB L2)"}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mgld", choices=["mgld", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--ddpm-steps", type=int, default=50)
    ap.add_argument("--frames", type=int, default=0, help="override the clip length (development)")
    ap.add_argument("--clips-per-batch", type=int, default=0, help="units sampled in lock-step as one (b t) batch (default: the pipeline's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the second (host-buffer) timed pass (long single-GPU runs of configs 4/5)")
    ap.add_argument("--flow", default="raft", choices=["raft", "synthetic"],
                    help="raft: RAFT_SR flow estimation inside the timed path (the reference's behaviour); synthetic: given flows")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 3 if args.config in (2, 3) else 1
    cfg = load_cfg()
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return
    args.warmup = max(args.warmup, 3)
    C = CONFIGS[args.config]
    n_frames = args.frames or C["frames"]
    seq_sharding = C["sharding"] == "sequence"

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
        shapes = mod.expected_shapes(False) if pre == "first_stage_model." else mod.expected_shapes()   # encoder half only
        sd.update({pre + k: v for k, v in fast_state_dict(shapes, seed).items()})
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
    if args.clips_per_batch > 0:
        pipe.clips_per_batch = args.clips_per_batch
        model.unet_clips_per_call = max(model.unet_clips_per_call, args.clips_per_batch)
    T = cfg.model.params.num_frames
    n_seg = (n_frames + T - 1) // T
    lr_h, lr_w = C["lr"]
    clip_host = synthetic_clip(42 + (rank if seq_sharding else 0), n_frames, lr_h, lr_w, C["degrade"]).pin_memory()
    H, W = 4 * lr_h, 4 * lr_w
    flows = None
    if args.flow == "synthetic":
        hp, wp = ((H // 32 + 1) * 32, (W // 32 + 1) * 32) if (H % 32 or W % 32) else (H, W)
        flows = [(a.to(dev), b.to(dev)) for a, b in synthetic_flows(7 + rank, n_seg, T, hp // 8, wp // 8)]
    out_host = torch.empty(n_frames, 3, H, W).pin_memory() if (rank == 0 or seq_sharding) else None
    gather_buf = torch.empty(world * n_frames, 3, H, W, dtype=torch.uint8, device=dev) if (world > 1 and seq_sharding) else None

    def run_clip(clip_dev):
        if seq_sharding:
            sr = pipe(clip_dev, context=context, flows_override=flows)
            if world > 1:   # stitch the sequences of all ranks: one all-gather of the finished uint8 frames
                dist.all_gather_into_tensor(gather_buf, (sr * 255.0).to(torch.uint8))   # truncation = script :529 astype(uint8)
            return sr
        return pipe(clip_dev, context=context, flows_override=flows, world_size=world, rank=rank)

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
    # warm-up: >= 3 passes.  Config 2/3: the whole clip.  Configs 4/5 (tens of seconds per pass): a prefix of the clip with
    # the same unit shapes, long enough that EVERY rank gets a full group of units.  Units are sampled in groups of
    # clips_per_batch plus one smaller remainder group (pipeline._sr_units), each group size a different (b t) batch shape
    # with its own CUDA graphs; which sizes a rank meets depends on its unit count, so the warm-up passes cycle through
    # the group sizes (clips_per_batch .. 1) and every graph the timed pass can need exists on every rank before it.
    units_per_segment = {4: 2, 5: 6}.get(args.config, 1)
    warm_segments = min(n_seg, max(2, -(-pipe.clips_per_batch * world // units_per_segment)))
    warm_dev = clip_dev if args.config in (2, 3) else clip_dev[:min(n_frames, warm_segments * T)]
    sizes = list(range(max(1, pipe.clips_per_batch), 0, -1))
    cpb = pipe.clips_per_batch
    for i in range(max(args.warmup, len(sizes)) if warm_dev is not clip_dev else args.warmup):
        if warm_dev is clip_dev:
            sr = run_clip(clip_dev)
        else:
            pipe.clips_per_batch = sizes[i % len(sizes)]
            sr = pipe(warm_dev, context=context, flows_override=None if flows is None else flows[:warm_segments],
                      world_size=world, rank=rank)
    pipe.clips_per_batch = cpb
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
        sr_ = run_clip(d)
        if out_host is not None:          # strong-scaled clip: every rank uploads the LR clip, rank 0 reads the result back
            out_host.copy_(sr_, non_blocking=True)
    if args.config in (2, 3) and not args.no_e2e:
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps) if not args.no_e2e else None
    sampler.stop_flag = True
    sampler.join(timeout=2)

    frames_total = n_frames * (world if seq_sharding else 1)
    value = frames_total * args.steps / (ms / 1e3)
    e2e_value = frames_total * args.steps / (ms_e2e / 1e3) if ms_e2e else None
    line = {"metric": C["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if seq_sharding else "strong",
            "vs_baseline": None, "dtype": "f16 (fp32 accumulate; fp32 norms/softmax/schedule/guidance)", "data": "synthetic",
            "config": workload_config(args, world), "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": clip_host.numel() * 4,
                    "d2h_bytes_per_step": n_frames * 3 * H * W * 4, "ms_per_step": ms_e2e / args.steps if ms_e2e else None},
            "gpu_launches": launches}
    if rank == 0:
        # ---- the other half of the BASELINE metric + rooflines of the two tensor-core kernels ------------------------------
        Tb = T * min(pipe.clips_per_batch, n_seg if args.config in (2, 3) else 2)
        line["unet_step_ms"] = measure_unet_step(model, dev, context, T, Tb)
        line["roofline"], line["roofline_attention"] = measure_rooflines(model, ops, dev, context, Tb)
        if world == 1 and args.config == 2:
            if not args.no_gpu_reference:
                line["gpu_reference"] = gpu_reference_sample(cfg, dev)
                line["gpu_reference"]["speedup_vs_gpu_reference"] = value / line["gpu_reference"]["value"]
            if not args.no_cpu_baseline:
                line["cpu_baseline"], _ = cpu_reference_sample(cfg, steps=1, warmup=0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_unet_step(model, dev, context, T, Tb):
    """"UNet step ms" (BASELINE metric): one struct-cond encoder + UNet evaluation on a (T=5,4,64,64) latent
    tile, as the sampler runs it (CUDA-graph replay), and on the (b t) = Tb-frame batch the timed workload uses per DDPM step."""
    out = {}
    for frames in sorted({T, Tb}):
        x, lat = torch.randn(frames, 4, 64, 64, device=dev), torch.randn(frames, 4, 64, 64, device=dev)
        t = torch.tensor([500], device=dev)
        for _ in range(3):
            model._eps(x, lat, t, context)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            model._eps(x, lat, t, context)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out[f"frames_{frames}"] = {"ms": ms, "ms_per_frame": ms / frames,
                                   "tflops": FLOP_TILE_STEP * frames / 5 / (ms * 1e-3) / 1e12}
    return out


def _peak():
    peak, src = 1590.0, "fallback 1.59 PFLOP/s (B200_PROFILING.md)"
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        try:
            j = json.load(open(pk))
            peak, src = float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "MEASURED_PEAKS.json bf16_tflops_sustained"
        except Exception:
            pass
    return peak, src


def measure_rooflines(model, ops, dev, context, T):
    """Every mgld_conv_gemm and mgld_attention launch of one eager struct-encoder + UNet tile-step (T frames = the `(b t)`
    batch the timed workload runs per DDPM step) is bracketed by CUDA events on the launching stream; achieved = sum of the
    algorithmic FLOPs / sum of the durations.  Peak: MEASURED_PEAKS.json (sustained: the kernels run inside a long step),
    else the B200_PROFILING.md fallback."""
    peak, src = _peak()
    rec, rec_a = [], []
    orig, orig_a = ops.conv_gemm, ops.attention

    def wrapped(a, w, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(a, w, **kw)
        e1.record()
        m = a.numel() // a.shape[-1]
        rec.append((e0, e1, 2.0 * m * w.shape[0] * w.shape[1]))
        return out

    def wrapped_a(q, k, v, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_a(q, k, v, **kw)
        e1.record()
        rec_a.append((e0, e1, 4.0 * kw["batch"] * kw["heads"] * kw["nq"] * kw["nkv"] * kw["head_dim"], kw["nq"], kw["nkv"]))
        return out

    x, lat = torch.randn(T, 4, 64, 64, device=dev), torch.randn(T, 4, 64, 64, device=dev)
    t = torch.tensor([500], device=dev)
    ops.conv_gemm, ops.attention = wrapped, wrapped_a
    try:
        for _ in range(2):
            rec.clear()
            rec_a.clear()
            model.model.diffusion_model(x, t, context=context, struct_cond=model.structcond_stage_model(lat, t))
        torch.cuda.synchronize()
    finally:
        ops.conv_gemm, ops.attention = orig, orig_a
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
    conv = {"kernel": "mgld::conv_gemm_kernel (tcgen05 implicit-GEMM conv/linear)", "bound": "tensor", "achieved": achieved,
            "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "peak_source": src,
            "launches_timed": len(rec), "flop_per_tile_step_in_kernel": tot_fl, "ms_per_tile_step_in_kernel": tot_ms,
            "frames_per_tile_step": T,
            "note": "achieved = sum of algorithmic FLOPs of the %d conv_gemm launches of one struct-enc+UNet tile-step / "
                    "sum of their CUDA-event durations (eager launches: includes ~2 us of event/launch gap each)" % len(rec)}
    a_ms = sum(e0.elapsed_time(e1) for e0, e1, *_ in rec_a)
    a_fl = sum(r[2] for r in rec_a)
    big = [r for r in rec_a if r[3] >= 4096 and r[4] >= 4096]
    attn = {"kernel": "all mgld_attention launches of the tile-step: mgld::attention_v3_kernel (tcgen05 flash attention, d=64), "
                      "cross_attention_kv80_kernel (77 text tokens), attention_kernel (small maps / d=128)", "bound": "tensor",
            "achieved": a_fl / (a_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": a_fl / (a_ms * 1e-3) / 1e12 / peak,
            "traffic": None, "peak_source": src, "launches_timed": len(rec_a), "flop_per_tile_step_in_kernel": a_fl,
            "ms_per_tile_step_in_kernel": a_ms, "frames_per_tile_step": T}
    if big:
        b_ms, b_fl = sum(r[0].elapsed_time(r[1]) for r in big), sum(r[2] for r in big)
        attn["self_attention_64x64"] = {"launches": len(big), "us_per_launch": 1e3 * b_ms / len(big),
                                        "tflops": b_fl / (b_ms * 1e-3) / 1e12, "frac": b_fl / (b_ms * 1e-3) / 1e12 / peak}
    return conv, attn


if __name__ == "__main__":
    main()
