#!/usr/bin/env python
"""Measured parity of the product against the TF32-free fp32 oracle, stage by stage, written to a JSON file
(committed as profiles/parity_r02.json).  Runs on the GPU box:

  python tests/parity_report.py [--out gpurun_out/parity.json] [--study]

Configurations: BASELINE configs[1] = C2 (batch 4, 512x512, 50 steps, 8-char strings) with per-step guided-eps errors
(free-running and teacher-forced on the oracle's x), C1 (golden of the unmodified reference's CPU run), tiny.
`--study` additionally evaluates the oracle itself with fp16-rounded weights / GEMM inputs / residual stream to show where
the fp16 floor of the path lies (which part of the error is inherent to fp16 storage, which part is the kernels').
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # parity_util

import torch
import torch.nn.functional as F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity.json"))
    ap.add_argument("--study", action="store_true")
    ap.add_argument("--study2", action="store_true", help="per-component fp16 weight rounding + oracle conditioning fed to the product")
    ap.add_argument("--skip-c1", action="store_true")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--batch", type=int, default=4)
    args = ap.parse_args()

    import parity_util as P
    from oracle import restated as R
    from udifftext_b200 import api, synth

    dev = torch.device("cuda", 0)
    report = {"oracle": "oracle/restated.py fp32 on the same B200, TF32 disabled (cudnn.allow_tf32 = matmul.allow_tf32 = False)",
              "metric": "rel = ||a-b||_2 / ||b||_2 over the whole tensor; maxabs = max |a-b|", "torch": torch.__version__}
    sd = synth.synthetic_state_dict(synth.load_manifest("full"), 1234)
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    eng = api.build_engine("full", dev, state_dict=sd)

    # ---------------------------------------------------------------- C2: BASELINE configs[1]
    steps, b = args.steps, args.batch
    check = sorted({0, 1, steps // 4, steps // 2, 3 * steps // 4, steps - 1})
    host = synth.synthetic_batch(2, b, 512, 512, 8)
    batch_dev = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    t0 = time.time()
    ora = P.oracle_run(R, sd_dev, batch_dev, steps, 5.0, 1002)
    torch.cuda.synchronize()
    t_or = time.time() - t0
    prod = P.product_run(api, eng, host, steps, 5.0, 1002, teacher_xs=ora["xs"], teacher_steps=check)
    c2 = P.compare(prod, ora, list(range(steps)))
    c2["eps_rel_checked_steps"] = {i: c2["eps_rel"][i] for i in check}
    c2["oracle_seconds"] = t_or
    # decoder alone: product decode of the ORACLE's latents
    with torch.no_grad():
        pix = eng.decode_first_stage_clamped(ora["z"].contiguous())
    c2["decoder_only_pixels_rel"] = P.rel(pix, ora["pixels"])
    c2["decoder_only_pixels_maxabs"] = P.maxabs(pix, ora["pixels"])
    c2["eps_rms_by_step"] = {i: float(ora["eps"][i].pow(2).mean().sqrt()) for i in check}
    report[f"C2_batch{b}_512px_{steps}steps"] = c2
    print(json.dumps({k: v for k, v in c2.items() if k not in ("eps_rel", "x_rel")}, indent=1), flush=True)

    if args.study:
        study = {}
        for mode in ("w", "wa", "was"):
            with P.quantised_oracle(R, mode):
                oq = P.oracle_run(R, sd_dev, batch_dev, steps, 5.0, 1002)
            study[mode] = {"pixels_rel": P.rel(oq["pixels"], ora["pixels"]), "latents_rel": P.rel(oq["z"], ora["z"]),
                           "eps_rel_step0": P.rel(oq["eps"][0], ora["eps"][0]),
                           "eps_rel_mid": P.rel(oq["eps"][steps // 2], ora["eps"][steps // 2]),
                           "eps_rel_last": P.rel(oq["eps"][steps - 1], ora["eps"][steps - 1]),
                           "product_vs_this_pixels_rel": P.rel(prod["pixels"], oq["pixels"])}
            print("study", mode, study[mode], flush=True)
        # TF32 oracle (what round 1 compared against)
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        b2, lat = b, (b, 4, 64, 64)
        torch.manual_seed(1002)
        noise_c, noise_uc, x0 = torch.randn(lat).to(dev), torch.randn(lat).to(dev), torch.randn(lat).to(dev)
        with torch.no_grad():
            c, uc = R.conditioner(R._sub(sd_dev, "conditioner."), batch_dev, noise_c, noise_uc)
            z = R.euler_sample(R._sub(sd_dev, "model.diffusion_model."), x0, c, uc, steps, 5.0)
            img = torch.clamp((R.vae_decode(R._sub(sd_dev, "first_stage_model."), z / 0.18215) + 1) / 2, 0, 1)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        study["tf32_oracle"] = {"pixels_rel": P.rel(img, ora["pixels"]), "latents_rel": P.rel(z, ora["z"])}
        print("study tf32", study["tf32_oracle"], flush=True)
        report["fp16_floor_study_C2"] = study

    if args.study2:
        st2 = {}
        for phase in ("cond", "unet", "dec"):
            for mode in ("w", "was"):
                oq = P.oracle_run(R, sd_dev, batch_dev, steps, 5.0, 1002, phase_ctx={phase: (lambda m=mode: P.quantised_oracle(R, m))})
                st2[f"{phase}_{mode}"] = {"pixels_rel": P.rel(oq["pixels"], ora["pixels"]), "latents_rel": P.rel(oq["z"], ora["z"]),
                                          "c_concat_rel": P.rel(oq["c"]["concat"], ora["c"]["concat"]),
                                          "t_crossattn_rel": P.rel(oq["c"]["t_crossattn"], ora["c"]["t_crossattn"])}
                print("study2", phase, mode, st2[f"{phase}_{mode}"], flush=True)
        pr2 = P.product_run(api, eng, host, steps, 5.0, 1002, cond_override=(ora["c"], ora["uc"]))
        st2["product_with_oracle_conditioning"] = {"latents_rel": P.rel(pr2["z"], ora["z"]), "pixels_rel": P.rel(pr2["pixels"], ora["pixels"])}
        print("study2 product with oracle conditioning", st2["product_with_oracle_conditioning"], flush=True)
        for key in ("t_crossattn", "concat"):
            c_mix = dict(prod["c"]); uc_mix = dict(prod["uc"])
            c_mix[key] = ora["c"][key]; uc_mix[key] = ora["uc"][key]
            pr3 = P.product_run(api, eng, host, steps, 5.0, 1002, cond_override=(c_mix, uc_mix))
            st2[f"product_with_oracle_{key}"] = {"latents_rel": P.rel(pr3["z"], ora["z"]), "pixels_rel": P.rel(pr3["pixels"], ora["pixels"])}
            print("study2 product with oracle", key, st2[f"product_with_oracle_{key}"], flush=True)
        report["component_study_C2"] = st2

    # ---------------------------------------------------------------- C1 vs the unmodified reference's CPU run (golden)
    if not args.skip_c1:
        gold = torch.load(os.path.join(ROOT, "tests", "golden", "c1.pt"))
        cfgs = api.runtime_config(steps=gold["steps"], batch_size=1, scale=[gold["scale"], 0.0])
        sampler = api.init_sampling(cfgs)
        sampler.verbose = False
        torch.manual_seed(gold["seed"])
        img, z = api.predict(cfgs, eng, sampler, synth.synthetic_batch(gold["config_id"], 1, 512, 512, 4))
        torch.cuda.synchronize()
        report["C1_1x512px_10steps_vs_reference_cpu_golden"] = {
            "latents_rel": P.rel(z.cpu(), gold["z"]), "pixels_rel": P.rel(img.cpu(), gold["pixels_f16"].float()),
            "pixels_maxabs": P.maxabs(img.cpu(), gold["pixels_f16"].float()),
            "note": "golden pixels are stored as fp16 (quantisation 2.4e-4 relative at most)"}
        print(report["C1_1x512px_10steps_vs_reference_cpu_golden"], flush=True)

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
