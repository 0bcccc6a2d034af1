"""Host-visible phases of one predict() (tuning aid): conditioner / init noise / sampler loop / decode, each bracketed by
torch.cuda.synchronize(), plus the same request un-instrumented."""
import sys
import time

import torch

sys.path.insert(0, ".")
from udifftext_b200 import api, synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    model = api.build_engine("full", dev)
    cfgs = api.runtime_config(steps=50, batch_size=B, gpu=0, noise_iters=0)
    sampler = api.init_sampling(cfgs)
    sampler.verbose = False
    batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.synthetic_batch(2, B, 512, 512, 8).items()}
    for i in range(3):
        torch.manual_seed(i)
        api.predict(cfgs, model, sampler, dict(batch))
    torch.cuda.synchronize()

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for rep in range(2):
        torch.manual_seed(10 + rep)
        with torch.no_grad():
            t0 = sync()
            b, buc = api.prepare_batch(cfgs, dict(batch))
            c, uc = model.conditioner.get_unconditional_conditioning(b, batch_uc=buc, force_uc_zero_embeddings=cfgs.force_uc_zero_embeddings)
            t1 = sync()
            x = sampler.get_init_noise(cfgs, model, cond=c, batch=b, uc=uc)
            t2 = sync()
            z = sampler(model, x, cond=c, batch=b, uc=uc, init_step=0, aae_enabled=False, detailed=False)
            t3 = sync()
            img = model.decode_first_stage_clamped(z)
            t4 = sync()
        print(f"rep {rep}: conditioner {1e3 * (t1 - t0):.2f} ms | init noise {1e3 * (t2 - t1):.2f} | sampler {1e3 * (t3 - t2):.2f} "
              f"| decode {1e3 * (t4 - t3):.2f} | total {1e3 * (t4 - t0):.2f}")
    t0 = sync()
    for rep in range(3):
        torch.manual_seed(20 + rep)
        api.predict(cfgs, model, sampler, dict(batch))
    t1 = sync()
    print(f"un-instrumented predict: {1e3 * (t1 - t0) / 3:.2f} ms per request of {B}")
    # python-side cost of the conditioner / decoder launches: time without sync between (host enqueue time)
    t0 = time.perf_counter()
    with torch.no_grad():
        b, buc = api.prepare_batch(cfgs, dict(batch))
        c, uc = model.conditioner.get_unconditional_conditioning(b, batch_uc=buc, force_uc_zero_embeddings=cfgs.force_uc_zero_embeddings)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"conditioner host enqueue time {1e3 * (t1 - t0):.2f} ms")


if __name__ == "__main__":
    main()
