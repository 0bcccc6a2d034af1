#!/usr/bin/env python
"""bench.py — 512x512, 50-step DDIM images/sec of the UDiffText inference hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]              our arm: hand-written sm_100a kernels
  python bench.py --impl reference [--gpus N --steps K --warmup W]  the reference's CPU path (oracle port) on the host cores
  torchrun ... bench.py --gpus N ...                               one rank per GPU, batch sharded, one NCCL all-gather

A "step" is one full `predict()` (test.py:19-40) over one batch: conditioner (LabelEncoder, mask rescale, VAE encode,
posterior sample) -> 50 CFG-doubled UNet + Euler steps -> VAE decode -> clamp.  Workload = BASELINE.json configs[1]:
batch 4 per GPU, 512x512, 50 steps, 8-character strings, synthetic inputs, seeded random weights of the exact
architecture (no checkpoints / datasets exist offline).  `value` times the path with the request tensors already
resident in HBM; `e2e` times the same call from pinned host buffers to host images (H2D + D2H inside the timed
region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

IMG = 512
DDIM_STEPS = 50
LABEL_LEN = 8
PER_GPU_BATCH = 4
# SURVEY.md §8(d) / BASELINE.md §2: algorithmic FLOPs measured on the unmodified reference module graph
GFLOP_UNET_SAMPLE_FWD = 798.66
GFLOP_UNET_IGEMM = 394.78 + 5.66 + 18.04 + 257.14   # conv3x3 + conv3x3-s2 + conv1x1 + linear: the udt_igemm share
GFLOP_VAE_ENC, GFLOP_VAE_DEC = 1116.7, 2514.5


def workload_name(batch: int, world: int = 1) -> str:
    """the `config.workload` string of both arms (identical text, so that the two lines can be matched)"""
    tag = ""
    if world == 1:
        tag = {4: " (BASELINE configs[1])", 32: " (BASELINE configs[2])"}.get(batch, "")
    elif batch == 4:
        tag = " (BASELINE configs[1] per GPU)"
    elif (batch, world) == (8, 8):
        tag = " (BASELINE configs[3])"
    return (f"batch {batch} per GPU, {IMG}x{IMG}, {DDIM_STEPS} DDIM steps (EulerEDM+LegacyDDPM, CFG 5.0), "
            f"{LABEL_LEN}-char strings, noise_iters 0{tag}")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))),
                "tflops_burst": float(p.get("bf16_tflops", 1590.0)), "hbm_gbs": float(p.get("hbm_gbs", 6650.0)),
                "source": "MEASURED_PEAKS.json"}
    return {"tflops": 1590.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md: 1.59 PFLOP/s, 6.65 TB/s)"}


def ncu_traffic_per_launch(batch: int = 4):
    """dram__bytes_read.sum + dram__bytes_write.sum per udt_igemm launch (average over the launches of one UNet CFG step),
    from the committed ncu capture of scripts/ncu_step.py at that batch size (profiles/r01_step_igemm_traffic.json for
    BASELINE configs[1], profiles/r02_step_igemm_traffic_b32.json for configs[2]); None for any other batch"""
    name = {4: "r01_step_igemm_traffic.json", 32: "r02_step_igemm_traffic_b32.json"}.get(batch)
    if name is None:
        return None
    path = os.path.join(ROOT, "profiles", name)
    try:
        with open(path) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)"""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[2 + j].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons, "samples": len(rows)}


# =================================================================================================== reference arm
def cpu_reference_sample(threads: int):
    """The reference's algorithm for this path on the host cores: oracle/restated.py (the fp32 PyTorch restatement
    pinned against the unmodified reference by oracle/make_golden.py; /root/reference itself does not exist on the
    GPU box).  Bounded sample of the workload: for ONE 512x512 image the conditioner's VAE encode, ONE CFG-doubled
    UNet evaluation and the VAE decode are timed; images/s for 50 steps = 1 / (t_enc + 50 t_unet + t_dec)
    (the reference runs the encoder twice per image — c and uc — so 2 t_enc is charged)."""
    import torch
    from oracle import restated as R
    from udifftext_b200 import synth
    torch.set_num_threads(threads)
    man = synth.load_manifest("full")
    sd = synth.synthetic_state_dict(man, 1234)
    unet_sd = R._sub(sd, "model.diffusion_model.")
    enc_sd = R._sub(sd, "conditioner.embedders.2.model.")
    dec_sd = R._sub(sd, "first_stage_model.")
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        x = torch.randn((2, 9, IMG // 8, IMG // 8), generator=g)
        ctx = torch.randn((2, 12, 2048), generator=g)
        t0 = time.perf_counter()
        R.unet_forward(unet_sd, x, torch.tensor([999, 999]), ctx)
        t_unet = time.perf_counter() - t0
        img = torch.rand((1, 3, IMG, IMG), generator=g) * 2 - 1
        t0 = time.perf_counter()
        R.vae_encode_moments(enc_sd, img)
        t_enc = time.perf_counter() - t0
        z = torch.randn((1, 4, IMG // 8, IMG // 8), generator=g)
        t0 = time.perf_counter()
        R.vae_decode(dec_sd, z)
        t_dec = time.perf_counter() - t0
    per_image = 2 * t_enc + DDIM_STEPS * t_unet + t_dec
    return {"value": 1.0 / per_image, "t_unet_cfg_s": t_unet, "t_enc_s": t_enc, "t_dec_s": t_dec, "seconds_per_image": per_image}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    for _ in range(args.warmup + args.steps):
        vals.append(cpu_reference_sample(threads))
    timed = vals[args.warmup:] if len(vals) > args.warmup else vals
    v = sum(t["value"] for t in timed) / len(timed)
    sample = ("1 image 512x512: VAE encode (x2) + one CFG UNet evaluation (batch 2) + VAE decode timed on the host; "
              "images/s = 1/(2 t_enc + 50 t_unet + t_dec)")
    line = {"impl": "reference", "metric": "512x512 50-step DDIM images/sec", "value": v, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * PER_GPU_BATCH / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(PER_GPU_BATCH, max(1, args.gpus)),
                       "note": "reference algorithm (fp32 PyTorch restatement, oracle/restated.py) on host CPU cores"},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample,
                             "detail": timed[-1]},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    args.emit(json.dumps(line))


def executed_igemm(prof: dict, ms_used: float, peak: float):
    """FLOPs the igemm launches of one step really EXECUTE (2 M N K of every launch as issued: zero-padded K blocks
    included, the four-phase upsample convs and the folded / skipped t_attn of the unconditional half save work against the
    reference's module graph) — the `achieved` figure above uses the reference's ALGORITHMIC FLOPs (effective throughput);
    this one says how busy the tensor pipe has to be for it"""
    flops = 0.0
    for r in prof.get("by_shape", []):
        if r["op"] != "udt_igemm":
            continue
        try:
            m, n, k = (float(v) for v in r["shape"][:3])
        except (ValueError, IndexError):
            continue
        flops += 2.0 * m * n * k * r["calls"]
    if flops == 0.0 or ms_used <= 0:
        return None
    t = flops / (ms_used * 1e-3) / 1e12
    return {"flop_per_unet_step": flops, "achieved": t, "frac": t / peak}


def gpu_library_baseline(dev, batch: int):
    """SURVEY.md §2.1 / BASELINE.md §4.2's bar: the reference's module graph executed by the GPU LIBRARIES (cuDNN convs,
    cuBLAS linears, SDPA attention; torch eager) on the same B200 — oracle/restated.py, the restatement pinned against the
    unmodified reference, because /root/reference cannot travel to this box.  One full predict of the same workload per
    precision: fp32 with torch's default flags (cuDNN TF32 on, cuBLAS fp32 — what the unmodified reference would run with)
    and torch.autocast(float16).  Baseline only: never on the product path."""
    import torch
    from oracle import restated as R
    from udifftext_b200 import synth
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = True, False      # torch defaults
    sd = {k: v.to(dev) for k, v in synth.synthetic_state_dict(synth.load_manifest("full"), 1234).items()}
    batch_dev = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v)
                 for k, v in synth.synthetic_batch(2, batch, IMG, IMG, LABEL_LEN).items()}
    out = {"what": "oracle/restated.py (reference module graph on cuDNN / cuBLAS / SDPA, torch eager) on the same B200, "
                   f"one predict of batch {batch}, {IMG}x{IMG}, {DDIM_STEPS} steps", "unit": "images/s"}

    def run(steps):
        torch.manual_seed(7)
        with torch.no_grad():
            return R.predict(sd, batch_dev, steps, 5.0)

    try:
        for name, ctx in (("fp32_cudnn_tf32", None), ("autocast_fp16", torch.autocast("cuda", dtype=torch.float16))):
            with (ctx if ctx is not None else torch.autocast("cuda", enabled=False)):
                run(2)                                    # warm-up: cuDNN heuristics, allocator
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run(DDIM_STEPS)
                e1.record()
                torch.cuda.synchronize(dev)
            out[name] = {"value": batch / (e0.elapsed_time(e1) * 1e-3), "ms_per_request": e0.elapsed_time(e1)}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
        del sd
        torch.cuda.empty_cache()
    return out


def pin_to_gpu_numa(local: int) -> str:
    """bind this rank's host threads to the CPU cores nearest its GPU (NVML's ideal CPU affinity), so that 8 ranks do not
    share one socket's cores while they launch kernels"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cores near GPU {local}"
    except Exception as e:      # noqa: BLE001 — affinity is an optimisation, never a failure
        return f"unchanged ({type(e).__name__})"
    return "unchanged"


# =================================================================================================== our arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from udifftext_b200 import api, ops, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — udifftext_b200 has no CPU path")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    affinity = pin_to_gpu_numa(local) if world > 1 else "not pinned (single rank)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = api.build_engine("full", dev)

    def measure(B: int, with_profile: bool):
        """device-timed and end-to-end images/s of `args.steps` requests of B images per GPU"""
        gb = B * world
        cfgs = api.runtime_config(steps=DDIM_STEPS, batch_size=B, gpu=local, noise_iters=0)
        sampler = api.init_sampling(cfgs)
        sampler.verbose = False
        # global synthetic request (seeded), this rank's rows; pinned host copies for the e2e leg
        full = synth.synthetic_batch(2, gb, IMG, IMG, LABEL_LEN)
        lo, hi = api.shard_bounds(gb, rank, world)
        host_batch = {k: (v.contiguous().pin_memory() if isinstance(v, torch.Tensor) else v)
                      for k, v in api.shard_batch(full, lo, hi).items()}
        dev_batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host_batch.items()}
        h2d = sum(v.numel() * v.element_size() for v in host_batch.values() if isinstance(v, torch.Tensor))
        # result of a request = the decoded images as uint8 HWC (demo.py:100-101 / test.py:94), converted on the device so
        # that NVLink (all-gather) and PCIe (D2H) carry a quarter of the fp32 bytes; only rank 0 reads the request back
        host_out = torch.empty((gb, IMG, IMG, 3), dtype=torch.uint8).pin_memory() if rank == 0 else None
        d2h = gb * IMG * IMG * 3

        def one_request(batch, seed, to_host: bool):
            torch.manual_seed(seed)
            img, _ = api.predict(cfgs, model, sampler, dict(batch), shard=(gb, lo, hi))
            u8 = api.images_to_u8(img)
            if world > 1:
                u8 = api.all_gather_images(u8, gb)   # the path's one collective: ncclAllGather of decoded images
            if to_host and rank == 0:
                host_out.copy_(u8, non_blocking=True)
            return u8

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        def timed(batch, to_host: bool):
            for i in range(args.warmup):
                one_request(batch, 100 + i, to_host)
            barrier()
            c0 = ops.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            clocks = ClockSampler(local)
            if rank == 0:
                clocks.start()
            e0.record()
            for i in range(args.steps):
                one_request(batch, 200 + i, to_host)
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            ck = clocks.stop() if rank == 0 else None
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms, ck, ops.launch_count() - c0

        ms_dev, clocks, eager_calls = timed(dev_batch, to_host=False)
        ms_e2e, clocks_e2e, _ = timed(host_batch, to_host=True)
        runner = sampler.last_runner
        # kernels launched in the timed region: graph replays (captured launches per step) + eager conditioner/decoder calls
        launches = eager_calls + args.steps * DDIM_STEPS * max(runner.launches_per_step, 1)
        imgs = gb * args.steps
        res = {"B": B, "gb": gb, "value": imgs / (ms_dev * 1e-3), "ms_per_step": ms_dev / args.steps,
               "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h if rank == 0 else 0, "ms_per_step": ms_e2e / args.steps,
                       "result": "uint8 HWC images (udt_images_to_u8 on the device), read back by rank 0 only"},
               "launches": int(launches), "clocks": clocks, "clocks_e2e": clocks_e2e}
        if with_profile:
            res["prof"] = ops.profile_step(runner, warm=2, reps=20)
            try:
                res["prof_graph"] = ops.profile_step_in_graph(runner)
            except Exception as e:      # noqa: BLE001 — the in-graph breakdown is evidence, not the product
                res["prof_graph"] = {"error": f"{type(e).__name__}: {e}"}
        return res

    B = args.batch
    m = measure(B, with_profile=True)
    prof, profg = m["prof"], m["prof_graph"]
    peaks = measured_peaks()
    ig = prof["by_op"].get("udt_igemm", {"ms": 0.0, "calls": 0})
    unet_step_ms = prof["step_ms_graph"]
    roofline = None
    if ig["ms"] > 0:
        flops = 2 * B * GFLOP_UNET_IGEMM * 1e9            # algorithmic FLOPs of all igemm launches of one CFG step
        alone = flops / (ig["ms"] * 1e-3) / 1e12
        igg = profg.get("by_op", {}).get("udt_igemm") if isinstance(profg, dict) else None
        if igg:      # primary: the kernel timed INSIDE the step graph (event-record nodes), against the sustained peak
            ach, peak, how = flops / (igg["ms"] * 1e-3) / 1e12, peaks["tflops"], \
                "in the step graph (external-event nodes between calls, 12 back-to-back replays) vs bf16_tflops_sustained"
            ms_used = igg["ms"]
        else:        # fallback: each shape replayed alone -> burst clocks, L2-warm: compare with the burst peak
            ach, peak, how, ms_used = alone, peaks["tflops_burst"], "each shape replayed alone vs bf16_tflops (burst)", ig["ms"]
        roofline = {"bound": "tensor", "kernel": "udt_igemm_kernel", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak, "traffic": ncu_traffic_per_launch(B), "peak_source": peaks["source"], "method": how,
                    "launches_per_unet_step": ig["calls"], "ms_per_unet_step": ms_used,
                    "avg_launch_us": 1e3 * ms_used / max(ig["calls"], 1), "flop_per_launch": flops / max(ig["calls"], 1),
                    "share_of_step": (igg["ms"] / profg["step_ms_graph_with_events"]) if igg else ig["ms"] / max(prof["step_ms_eager_sum"], 1e-9),
                    "alone": {"achieved": alone, "peak": peaks["tflops_burst"], "frac": alone / peaks["tflops_burst"],
                              "ms_per_unet_step": ig["ms"], "method": "each distinct shape replayed 20x alone in a CUDA graph (burst "
                              "clocks, small problems L2-warm) vs bf16_tflops (burst)"},
                    "executed": executed_igemm(prof, ms_used, peak),
                    "step_ms_graph": unet_step_ms,
                    "step_ms_graph_with_events": profg.get("step_ms_graph_with_events") if isinstance(profg, dict) else None}

    # BASELINE configs[3]: batch 64 sharded 8 per GPU over 8 GPUs — measured in the same run when this is the 8-rank job
    c3 = None
    if world == 8 and B != 8 and not args.no_configs3:
        r3 = measure(8, with_profile=False)
        c3 = {"workload": "BASELINE configs[3]: batch 64 sharded 8 per GPU across 8 GPUs, 50 steps, one NCCL all-gather of the decoded images",
              "value": r3["value"], "unit": "images/s", "ms_per_step": r3["ms_per_step"], "e2e": r3["e2e"], "global_batch": r3["gb"]}

    line = None
    if rank == 0:
        line = {"metric": "512x512 50-step DDIM images/sec", "value": m["value"], "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "fp16 (fp32 accumulate)", "data": "synthetic",
                "config": {"workload": workload_name(B, world),
                           "global_batch": m["gb"], "parallelism": f"batch-sharded x{world}, one all-gather of decoded uint8 images",
                           "l2": "working set per step >> 126 MB L2 (1.78 GB fp16 UNet weights streamed every step); no explicit flush",
                           "weights": "seeded synthetic, exact SD-2-inpainting UNifiedUNet / AutoencoderKL / LabelEncoder architecture",
                           "host_affinity": affinity},
                "unet_step_ms": unet_step_ms,
                "unet_step_tflops": 2 * B * GFLOP_UNET_SAMPLE_FWD / unet_step_ms if unet_step_ms else None,
                "e2e": m["e2e"],
                "gpu_launches": m["launches"],
                "clocks": m["clocks"], "clocks_e2e": m["clocks_e2e"],
                "roofline": roofline,
                "kernel_breakdown_unet_step": prof["by_op"],
                "kernel_breakdown_unet_step_in_graph": profg.get("by_op") if isinstance(profg, dict) else None,
                "top_calls_unet_step": prof["by_shape"][:12]}
        if c3 is not None:
            line["configs3"] = c3
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        try:
            line["gpu_library_baseline"] = gpu_library_baseline(dev, B)
        except Exception as e:      # noqa: BLE001
            line["gpu_library_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        s = cpu_reference_sample(threads)
        line["cpu_baseline"] = {"value": s["value"], "unit": "images/s", "cores": threads, "kind": "port",
                                "sample": "1 image 512x512 on the host cores via oracle/restated.py: VAE encode (x2) + one CFG UNet "
                                          "evaluation (batch 2) + VAE decode; images/s = 1/(2 t_enc + 50 t_unet + t_dec)",
                                "detail": s}
    if rank == 0:
        args.emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def claim_stdout():
    """Route everything that libraries print on fd 1 (e.g. NCCL's version banner) to stderr and return a writer for the
    ONE JSON line the contract allows on stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), "w", buffering=1)

    def emit(line: str) -> None:
        os.write(real, (line + "\n").encode())

    return emit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="images per GPU per request")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the GPU-library baseline (oracle on cuDNN/cuBLAS/SDPA)")
    ap.add_argument("--no-configs3", action="store_true", help="8-rank runs: skip the extra BASELINE configs[3] (8 per GPU) leg")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    relaunch = args.impl != "reference" and args.gpus != world and world == 1 and args.gpus > 1
    if not relaunch:
        args.emit = claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus != world and world == 1 and args.gpus > 1:
            # convenience: python bench.py --gpus N re-launches itself under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
        run_b200(args)


if __name__ == "__main__":
    main()
