"""End-to-end parity of `predict` (test.py:19-40) on the GPU against golden vectors produced by the UNMODIFIED
reference in the build container (oracle/make_golden.py): conditioner outputs (RNG-order check), final latents and
decoded pixels.  Tolerances: the product computes in fp16 storage / fp32 accumulation, the reference in fp32."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def tiny_engine(udt_lib):
    from udifftext_b200 import api
    return api.build_engine("tiny", torch.device("cuda", 0))


def test_tiny_conditioner_matches_reference(tiny_engine):
    from udifftext_b200 import api, synth
    gold = torch.load(os.path.join(GOLD, "tiny.pt"))
    cfgs = api.runtime_config(steps=gold["predict_steps"], batch_size=2, scale=[gold["predict_scale"], 0.0])
    batch = synth.synthetic_batch(gold["predict_config_id"], 2, 64, 64, None)
    torch.manual_seed(gold["predict_seed"])
    batch, batch_uc = api.prepare_batch(cfgs, batch)
    c, uc = tiny_engine.conditioner.get_unconditional_conditioning(batch, batch_uc=batch_uc, force_uc_zero_embeddings=["label"])
    torch.cuda.synchronize()
    # the c and uc latents are two different posterior draws of the same moments, in this order
    assert _rel(c["concat"], gold["c_concat"]) < 5e-3
    assert _rel(uc["concat"], gold["uc_concat"]) < 5e-3
    assert (c["concat"][:, 1:] - uc["concat"][:, 1:]).abs().max().item() > 1e-3
    assert torch.equal(c["concat"][:, :1], uc["concat"][:, :1])
    assert _rel(c["t_crossattn"], gold["c_crossattn"]) < 5e-3
    assert uc["t_crossattn"].abs().max().item() == 0.0
    # generic (unfused) conditioner path: same RNG stream, same values
    torch.manual_seed(gold["predict_seed"])
    c2, uc2 = tiny_engine.conditioner(batch), tiny_engine.conditioner(batch_uc, ["label"])
    assert _rel(c2["concat"], c["concat"]) < 1e-6 and _rel(uc2["concat"], uc["concat"]) < 1e-6


def test_tiny_predict_matches_reference(tiny_engine):
    from udifftext_b200 import api, synth
    gold = torch.load(os.path.join(GOLD, "tiny.pt"))
    cfgs = api.runtime_config(steps=gold["predict_steps"], batch_size=2, scale=[gold["predict_scale"], 0.0])
    sampler = api.init_sampling(cfgs)
    sampler.verbose = False
    batch = synth.synthetic_batch(gold["predict_config_id"], 2, 64, 64, None)
    torch.manual_seed(gold["predict_seed"])
    img, z = api.predict(cfgs, tiny_engine, sampler, batch)
    torch.cuda.synchronize()
    ez, ep = _rel(z, gold["predict_z"]), _rel(img, gold["predict_pixels"])
    print(f"tiny predict: z rel-L2 {ez:.3e}, pixels rel-L2 {ep:.3e}, max-abs {(img.cpu() - gold['predict_pixels']).abs().max():.3e}")
    assert ez < 2e-2 and ep < 2e-2
    # a second request replays the captured CUDA graph: bit-identical
    torch.manual_seed(gold["predict_seed"])
    img2, z2 = api.predict(cfgs, tiny_engine, sampler, synth.synthetic_batch(gold["predict_config_id"], 2, 64, 64, None))
    assert torch.equal(z, z2) and torch.equal(img, img2)
    # the generic per-op sampler path (sampler_step) gives the same latents as the fused graph
    torch.manual_seed(gold["predict_seed"])
    b, buc = api.prepare_batch(cfgs, synth.synthetic_batch(gold["predict_config_id"], 2, 64, 64, None))
    c, uc = tiny_engine.conditioner.get_unconditional_conditioning(b, batch_uc=buc, force_uc_zero_embeddings=["label"])
    x = sampler.get_init_noise(cfgs, tiny_engine, cond=c, batch=b, uc=uc)
    x, s_in, sigmas, n, c, uc = sampler.prepare_sampling_loop(x, c, uc)
    for i in range(n - 1):
        x, _, _ = sampler.sampler_step(s_in * sigmas[i], s_in * sigmas[i + 1], tiny_engine, x, c, b, uc)
    torch.cuda.synchronize()
    assert _rel(x, z) < 2e-3


def test_sharded_predict_is_rank_count_independent(tiny_engine):
    """rows [lo, hi) of a sharded request equal the same rows of the unsharded request (global RNG, sliced)"""
    from udifftext_b200 import api, synth
    cfgs = api.runtime_config(steps=3, batch_size=4)
    sampler = api.init_sampling(cfgs)
    sampler.verbose = False
    full = synth.synthetic_batch(7, 4, 64, 64, None)
    torch.manual_seed(5)
    img, z = api.predict(cfgs, tiny_engine, sampler, dict(full))
    cfg2 = api.runtime_config(steps=3, batch_size=2)
    parts = []
    for lo in (0, 2):
        sub = {k: (v[lo: lo + 2] if isinstance(v, (torch.Tensor, list)) else v) for k, v in full.items()}
        torch.manual_seed(5)
        parts.append(api.predict(cfg2, tiny_engine, sampler, sub, shard=(4, lo, lo + 2))[0])
    torch.cuda.synchronize()
    assert _rel(torch.cat(parts), img) < 1e-3


def test_c1_full_predict_matches_reference(udt_lib):
    """BASELINE.json configs[0]: 1x512x512, 4-char string, 10 steps — against the reference's CPU fp32 run"""
    from udifftext_b200 import api, synth
    gold = torch.load(os.path.join(GOLD, "c1.pt"))
    eng = api.build_engine("full", torch.device("cuda", 0))
    cfgs = api.runtime_config(steps=gold["steps"], batch_size=1, scale=[gold["scale"], 0.0])
    sampler = api.init_sampling(cfgs)
    sampler.verbose = False
    batch = synth.synthetic_batch(gold["config_id"], 1, 512, 512, 4)
    torch.manual_seed(gold["seed"])
    img, z = api.predict(cfgs, eng, sampler, batch)
    torch.cuda.synchronize()
    ez, ep = _rel(z, gold["z"]), _rel(img, gold["pixels_f16"])
    print(f"C1 predict: z rel-L2 {ez:.3e}, pixels rel-L2 {ep:.3e}, max-abs {(img.cpu() - gold['pixels_f16'].float()).abs().max():.3e}")
    assert ez < 3e-2 and ep < 3e-2
