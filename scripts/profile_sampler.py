import sys, time, torch
sys.path.insert(0, ".")
from udifftext_b200 import api, synth
dev = torch.device("cuda", 0)
B = 4
model = api.build_engine("full", dev)
cfgs = api.runtime_config(steps=50, batch_size=B, gpu=0, noise_iters=0)
sampler = api.init_sampling(cfgs); sampler.verbose = False
batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.synthetic_batch(2, B, 512, 512, 8).items()}
for i in range(2):
    torch.manual_seed(i); api.predict(cfgs, model, sampler, dict(batch))
torch.cuda.synchronize()
def sync():
    torch.cuda.synchronize(); return time.perf_counter()
with torch.no_grad():
    b, buc = api.prepare_batch(cfgs, dict(batch))
    c, uc = model.conditioner.get_unconditional_conditioning(b, batch_uc=buc, force_uc_zero_embeddings=cfgs.force_uc_zero_embeddings)
    x = sampler.get_init_noise(cfgs, model, cond=c, batch=b, uc=uc)
    for rep in range(3):
        t0 = sync()
        x2, s_in, sigmas, num_sigmas, c2, uc2 = sampler.prepare_sampling_loop(x.clone(), c, uc, None)
        t1 = sync()
        runner = sampler._runner(model, x2, c2)
        runner.begin(x2, c2, uc2, model.denoiser, sigmas, 0.0, 0.0, 999.0)
        t2 = sync()
        for i in range(num_sigmas - 1):
            runner.step(i)
        t3 = sync()
        t4h = time.perf_counter()
        for i in range(num_sigmas - 1):
            runner.replay()
        t4e = time.perf_counter()
        t4 = sync()
        print(f"prepare {1e3*(t1-t0):.2f} ms | begin {1e3*(t2-t1):.2f} | 50 steps {1e3*(t3-t2):.2f} | 50 bare replays {1e3*(t4-t3):.2f} (host enqueue {1e3*(t4e-t4h):.2f})")
