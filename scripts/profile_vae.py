"""Per-shape device time of the VAE decode / encode / LabelEncoder calls of one request (ops.profile_callable).  Tuning aid."""
import sys

import torch

sys.path.insert(0, ".")
from udifftext_b200 import api, ops, synth  # noqa: E402


def show(title, res, top=22):
    print(f"== {title}: {res['total_ms']:.3f} ms in C-ABI calls  " + "  ".join(f"{k}={v['ms']:.3f}" for k, v in sorted(res["by_op"].items(), key=lambda kv: -kv[1]["ms"])))
    for r in res["by_shape"][:top]:
        extra = ""
        if r["op"] == "udt_igemm":
            m, n, k = int(r["shape"][0]), int(r["shape"][1]), int(r["shape"][2])
            extra = f"  {2.0 * m * n * k / r['us_per_call'] / 1e6:7.0f} TFLOP/s"
        print(f"  {r['us_per_call'] * r['calls'] / 1e3:8.4f} ms  x{r['calls']:<3d} {r['us_per_call']:8.2f} us  {r['op']:22s} {tuple(r['shape'])}{extra}")


def main():
    dev = torch.device("cuda", 0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    model = api.build_engine("full", dev)
    cfgs = api.runtime_config(steps=50, batch_size=B, gpu=0, noise_iters=0)
    batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.synthetic_batch(2, B, 512, 512, 8).items()}
    z = torch.randn((B, 4, 64, 64), device=dev)
    with torch.no_grad():
        for _ in range(2):
            model.decode_first_stage_clamped(z)
        torch.cuda.synchronize()
        show(f"VAE decode, {B} images", ops.profile_callable(lambda: model.decode_first_stage_clamped(z)))
        b, buc = api.prepare_batch(cfgs, dict(batch))
        f = lambda: model.conditioner.get_unconditional_conditioning(b, batch_uc=buc, force_uc_zero_embeddings=cfgs.force_uc_zero_embeddings)
        f()
        torch.cuda.synchronize()
        show(f"conditioner (LabelEncoder + VAE encode), {B} images", ops.profile_callable(f))


if __name__ == "__main__":
    main()
