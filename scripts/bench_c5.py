"""BASELINE configs[4]: batch 8, 768x768 images (96x96 latents, 9216-token self-attention), 100 DDIM steps, 1 GPU.
Prints one JSON line (images/s, UNet CFG step ms).  usage: python scripts/bench_c5.py [batch=8] [steps=100] [px=768]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from udifftext_b200 import api, ops, synth  # noqa: E402


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    px = int(sys.argv[3]) if len(sys.argv) > 3 else 768
    dev = torch.device("cuda", 0)
    model = api.build_engine("full", dev)
    cfgs = api.runtime_config(steps=steps, batch_size=b)
    sampler = api.init_sampling(cfgs)
    sampler.verbose = False
    batch = synth.synthetic_batch(5, b, px, px, None)
    dbatch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    torch.manual_seed(1005)
    api.predict(cfgs, model, sampler, dict(dbatch))            # warm-up: builds the step graph for this shape
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.manual_seed(1005)
    e0.record()
    img, z = api.predict(cfgs, model, sampler, dict(dbatch))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    r = sampler.last_runner
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(5):
        r.replay()
    g1.record()
    torch.cuda.synchronize()
    print(json.dumps({"config": f"batch {b}, {px}x{px}, {steps} steps (BASELINE configs[4])", "images_per_s": b / ms * 1e3,
                      "ms_per_request": ms, "unet_step_ms_graph": g0.elapsed_time(g1) / 5, "launches_per_step": r.launches_per_step,
                      "c_abi_calls": ops.launch_count() - n0, "pixels_finite": bool(torch.isfinite(img).all()),
                      "pixel_mean": float(img.mean())}))


if __name__ == "__main__":
    main()
