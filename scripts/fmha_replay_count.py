"""How often the FMHA single pass is replayed / the running output rescaled on the REAL request (tuning build,
UDT_FMHA_DEBUG=32: every 37th CTA prints per-warp counts at its end).  usage: UDT_FMHA_DEBUG=32 python scripts/fmha_replay_count.py"""
import collections, os, re, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ".")
    from udifftext_b200 import api, synth
    dev = torch.device("cuda", 0)
    model = api.build_engine("full", dev)
    cfgs = api.runtime_config(steps=10, batch_size=4, gpu=0, noise_iters=0)
    sampler = api.init_sampling(cfgs)
    sampler.verbose = False
    batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synth.synthetic_batch(2, 4, 512, 512, 8).items()}
    torch.manual_seed(1)
    api.predict(cfgs, model, sampler, dict(batch))
    torch.cuda.synchronize()
    sys.exit(0)
env = dict(os.environ, UDT_FMHA_DEBUG="32")
out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, 0])
for m in re.finditer(r"FCNT nkv (\d+) heads (\d+) cta \d+ warp \d+ replays (\d+) rescales (\d+)", out):
    nkv, heads, rp, rs = (int(x) for x in m.groups())
    a = agg[(nkv, heads)]
    a[0] += 1; a[1] += rp; a[2] += rs
for (nkv, heads), (n, rp, rs) in sorted(agg.items()):
    print(f"key tiles {nkv:3d} heads {heads:2d}: {n} warp samples, replays per warp-item {rp / n:.2f} of {nkv - 1} single-pass tiles, rescales {rs / n:.2f}")
