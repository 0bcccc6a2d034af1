"""Container-only generator of the committed golden vectors (tests/golden/*.pt) — TEST INFRASTRUCTURE.

Runs the UNMODIFIED reference (imported from /root/reference with the shims of oracle/ref_import.py) on the
seeded synthetic weights / batches of udifftext_b200.synth, checks that oracle/restated.py reproduces it, and
stores the reference's outputs.  The GPU box cannot see the reference; its tests compare the CUDA path with
these vectors and with the restatement.

One deliberate deviation, documented in DESIGN.md: the reference leaves its LabelEncoder in *training* mode at
inference (`embedder.train = disabled_train` is installed before `model.eval()`, encoders/modules.py:117-119,
util.py:18-20), so dropout(p=0.1) makes the text embedding random.  Golden vectors are generated with the
LabelEncoder switched to eval mode (`nn.Module.train(le, False)`) — the deterministic function both sides share.

usage: python oracle/make_golden.py [tiny] [loss] [noise_search] [full_unet] [c1] [parseq]
"""
import os
import sys
import time

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, restated as R  # noqa: E402
from oracle.make_manifest import small_overrides  # noqa: E402
from udifftext_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SAMPLER_CFG = dict(
    discretization_config={"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"},
    s_churn=0.0, s_tmin=0.0, s_tmax=999.0, s_noise=1.0, verbose=False, device="cpu")


def reference_engine(name: str, seed: int = 1234):
    m = ref_import.build_reference_engine(seed, small_overrides(name))
    sd = synth.synthetic_state_dict(synth.load_manifest(name), seed)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    nn.Module.train(m.conditioner.embedders[0], False)  # see module docstring
    return m, sd


def reference_predict(m, batch, steps, scale, seed):
    """test.py:19-40 on CPU (the script hard-codes cuda; same calls, same RNG order, noise_iters = 0)."""
    from sgm.modules.diffusionmodules.sampling import EulerEDMSampler
    sampler = EulerEDMSampler(num_steps=steps, guider_config={
        "target": "sgm.modules.diffusionmodules.guiders.VanillaCFG", "params": {"scale": scale}}, **SAMPLER_CFG)
    batch_uc = dict(batch)
    batch_uc["txt"] = ["" for _ in batch["txt"]]
    batch_uc["label"] = ["" for _ in batch["label"]]
    torch.manual_seed(seed)
    with torch.no_grad():
        c, uc = m.conditioner.get_unconditional_conditioning(batch, batch_uc=batch_uc, force_uc_zero_embeddings=["label"])
        b, _, hh, ww = batch["image"].shape
        x = torch.randn((b, 4, hh // 8, ww // 8))
        z = sampler(m, x, cond=c, batch=batch, uc=uc, init_step=0, aae_enabled=False, detailed=False)
        img = torch.clamp((m.decode_first_stage(z) + 1.0) / 2.0, 0.0, 1.0)
    return img, z, c, uc


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def gen_tiny():
    m, sd = reference_engine("tiny")
    a = synth.ARCH["tiny"]
    g = torch.Generator().manual_seed(7)
    x = torch.randn((4, 9, 8, 8), generator=g)
    t = torch.tensor([999, 500, 20, 3])
    ctx = torch.randn((4, 12, a["unet"]["t_context_dim"]), generator=g)
    with torch.no_grad():
        ref = m.model.diffusion_model(x, timesteps=t, t_context=ctx)
        probs = [it["attn_map"].clone() for it in m.model.diffusion_model.attn_map_cache]
        mine_probs = []
        mine = R.unet_forward(R._sub(sd, "model.diffusion_model."), x, t, ctx, model_channels=a["unet"]["model_channels"],
                              probs_out=mine_probs)
    assert rel(mine, ref) < 1e-5, rel(mine, ref)
    assert max((p - q).abs().max().item() for p, q in zip(probs, mine_probs)) < 1e-5
    batch = synth.synthetic_batch(101, 2, 64, 64, None)
    img, z, c, uc = reference_predict(m, batch, 4, 5.0, 1101)
    torch.manual_seed(1101)
    with torch.no_grad():
        img2, z2 = R.predict(sd, batch, 4, 5.0)
    print("tiny: unet rel", rel(mine, ref), "predict z rel", rel(z2, z), "pixels rel", rel(img2, img), "eps rms",
          ref.pow(2).mean().sqrt().item(), "pixel mean", img.mean().item(), "sat frac",
          ((img == 0) | (img == 1)).float().mean().item())
    assert rel(z2, z) < 1e-4 and rel(img2, img) < 1e-4
    torch.save({"unet_x": x, "unet_t": t, "unet_ctx": ctx, "unet_out": ref, "unet_probs": probs,
                "predict_config_id": 101, "predict_seed": 1101, "predict_steps": 4, "predict_scale": 5.0,
                "predict_pixels": img, "predict_z": z, "c_concat": c["concat"], "uc_concat": uc["concat"],
                "c_crossattn": c["t_crossattn"]}, os.path.join(GOLD, "tiny.pt"))


def gen_full_unet():
    m, sd = reference_engine("full")
    g = torch.Generator().manual_seed(11)
    x = torch.randn((2, 9, 64, 64), generator=g)
    t = torch.tensor([999, 19])
    ctx = torch.randn((2, 12, 2048), generator=g)
    t0 = time.time()
    with torch.no_grad():
        ref = m.model.diffusion_model(x, timesteps=t, t_context=ctx)
        dt = time.time() - t0
        probs6 = [it["attn_map"][:, ::64].clone() for it in m.model.diffusion_model.attn_map_cache]
        mine = R.unet_forward(R._sub(sd, "model.diffusion_model."), x, t, ctx)
    print("full unet: rel", rel(mine, ref), "eps rms", ref.pow(2).mean().sqrt().item(), "ref seconds", dt)
    assert rel(mine, ref) < 1e-5
    torch.save({"x": x, "t": t, "ctx": ctx, "out": ref, "probs_strided64": probs6, "cpu_seconds_batch2": dt,
                "threads": torch.get_num_threads()}, os.path.join(GOLD, "full_unet.pt"))
    return m, sd


def gen_c1(m=None, sd=None):
    """BASELINE.json configs[0]: 1x512x512, 4-char string, 10 steps, reference PyTorch CPU fp32."""
    if m is None:
        m, sd = reference_engine("full")
    batch = synth.synthetic_batch(1, 1, 512, 512, 4)
    t0 = time.time()
    img, z, c, uc = reference_predict(m, batch, 10, 5.0, 1001)
    dt = time.time() - t0
    print("c1: reference predict seconds", dt, "pixel mean", img.mean().item(), "sat frac",
          ((img == 0) | (img == 1)).float().mean().item(), "z rms", z.pow(2).mean().sqrt().item())
    torch.save({"config_id": 1, "seed": 1001, "steps": 10, "scale": 5.0, "pixels_f16": img.half(), "z": z,
                "c_concat": c["concat"], "uc_concat": uc["concat"], "c_crossattn": c["t_crossattn"],
                "cpu_seconds": dt, "threads": torch.get_num_threads()}, os.path.join(GOLD, "c1.pt"))


def gen_loss():
    """the reference's FullLoss.get_min_local_loss (loss.py:192-235) on synthetic attention maps: the one-image CFG-doubled
    case the reference supports (sampling.py:340, [uc; c] maps against a [1, ...] mask), with the test.yaml Gaussian."""
    m, _ = reference_engine("tiny")
    lf = m.loss_fn
    g = torch.Generator().manual_seed(23)
    cache = []
    for name, heads, size in (("output_blocks.1.1.transformer_blocks.0.t_attn", 3, 16), ("mid.attn1", 2, 16),
                              ("input_blocks.1.1.transformer_blocks.0.t_attn", 2, 32),
                              ("input_blocks.7.1.transformer_blocks.0.t_attn", 4, 8)):
        probs = (torch.randn((2 * heads, size * size, 12), generator=g) * 2.0).softmax(-1)
        cache.append({"name": name, "heads": heads, "size": size, "attn_map": probs})
    mask = torch.zeros((1, 1, 128, 128))
    mask[:, :, 40:72, 20:100] = 1.0
    seg = torch.zeros((1, 12))
    seg[0, :5] = 1.0
    with torch.no_grad():
        ref = lf.get_min_local_loss(cache, mask, seg)
    k, sig = lf.gaussian_kernel_size, 1.0   # configs/test/textdesign_sd_2.yaml:113-114
    assert torch.allclose(R.gaussian_kernel(k, sig), lf.g_kernel[0, 0], atol=1e-7)
    mine = R.min_local_loss(cache, mask, seg, k, sig, lf.min_attn_size)
    print("loss: reference", ref.tolist(), "restated", mine.tolist(), "min_attn_size", lf.min_attn_size)
    assert ref.shape == (2,) and torch.allclose(mine, ref, atol=1e-7)
    torch.save({"cache": cache, "mask": mask, "seg_mask": seg, "kernel_size": k, "sigma": sig,
                "min_attn_size": lf.min_attn_size, "g_kernel": lf.g_kernel.clone(), "loss": ref},
               os.path.join(GOLD, "loss.pt"))


class _CpuDeviceTorch:
    """stand-in for the `torch` global of sgm/modules/diffusionmodules/sampling.py while get_init_noise runs: the function
    hard-codes torch.device("cuda", index=cfgs.gpu) (sampling.py:269,311); every other attribute is the real torch"""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*args, **kwargs):
        return torch.device("cpu")


def gen_noise_search():
    """the reference's EulerEDMSampler.get_init_noise (sampling.py:264-322) UNMODIFIED on CPU for one image (the case it
    supports), 3 trial noises; the per-trial scores are recorded by wrapping loss_fn.get_min_local_loss"""
    import types
    m, sd = reference_engine("tiny")                     # imports the reference (shims installed)
    import sgm.modules.diffusionmodules.sampling as S
    sampler = S.EulerEDMSampler(num_steps=4, guider_config={
        "target": "sgm.modules.diffusionmodules.guiders.VanillaCFG", "params": {"scale": 5.0}}, **SAMPLER_CFG)
    batch = synth.synthetic_batch(102, 1, 64, 64, 5)
    batch_uc = dict(batch, txt=[""], label=[""])
    iters = 3
    cfgs = types.SimpleNamespace(batch_size=1, channel=4, factor=8, noise_iters=iters, gpu=0)
    m.loss_fn.min_attn_size = 4          # configuration (loss_fn_config.params.min_attn_size): the tiny maps are 8x8 and 4x4
    scores = []
    orig = m.loss_fn.get_min_local_loss

    def recording(*a, **k):
        out = orig(*a, **k)
        scores.append(out.detach().clone())
        return out

    m.loss_fn.get_min_local_loss = recording
    real_torch = S.torch
    torch.manual_seed(2102)
    with torch.no_grad():
        c, uc = m.conditioner.get_unconditional_conditioning(batch, batch_uc=batch_uc, force_uc_zero_embeddings=["label"])
        state = torch.get_rng_state()
        S.torch = _CpuDeviceTorch()
        try:
            best = sampler.get_init_noise(cfgs, m, c, batch, uc)
        finally:
            S.torch = real_torch
            m.loss_fn.get_min_local_loss = orig
    assert len(scores) == 2 * iters                       # one score per sampler step, two steps per trial
    last = torch.stack([s_[s_.shape[0] // 2:] for s_ in scores[1::2]])          # [iters, 1]: conditional half, last step
    torch.set_rng_state(state)
    noises = [torch.randn((1, 4, 8, 8)) for _ in range(iters)]
    lf = m.loss_fn
    with torch.no_grad():
        mine_best, mine_losses = R.init_noise_search(R._sub(sd, "model.diffusion_model."), noises, c, uc, batch["mask"],
                                                     batch["seg_mask"], 5.0, lf.gaussian_kernel_size, 1.0, lf.min_attn_size)
    print("noise search: reference scores", last.flatten().tolist(), "restated", mine_losses.flatten().tolist())
    assert torch.allclose(mine_losses, last, atol=1e-5) and torch.equal(mine_best, best)
    torch.save({"config_id": 102, "label_len": 5, "seed": 2102, "iters": iters, "scale": 5.0, "noises": noises, "best": best,
                "losses": last, "c_concat": c["concat"], "uc_concat": uc["concat"], "c_crossattn": c["t_crossattn"],
                "uc_crossattn": uc["t_crossattn"], "kernel_size": lf.gaussian_kernel_size, "sigma": 1.0,
                "min_attn_size": lf.min_attn_size}, os.path.join(GOLD, "noise_search.pt"))


def gen_parseq():
    """PARSeq (OCR scoring): the unmodified reference PARSeq (AR decode + 1 refinement) on seeded weights and
    seeded 32x128 inputs -> logits + decoded strings; oracle/parseq_restated.py must reproduce both.  Also pins
    ParseqPredictor's crop preprocessing (torchvision Resize BICUBIC antialias + Normalize) against torchvision itself."""
    from oracle import parseq_restated as PR
    PARSeq = ref_import.import_reference_parseq()
    import yaml
    cfgd = os.path.join(ref_import.REFERENCE_ROOT, "src", "parseq", "configs")
    cfg = yaml.safe_load(open(os.path.join(cfgd, "main.yaml")))["model"]
    cfg.update(yaml.safe_load(open(os.path.join(cfgd, "charset", "94_full.yaml")))["model"])
    cfg.update(yaml.safe_load(open(os.path.join(cfgd, "model", "parseq.yaml"))))
    for k in ("_convert_", "_target_", "name"):
        cfg.pop(k, None)
    cfg["lr"] = float(cfg["lr"])
    model = PARSeq(**cfg).eval()
    sd = synth.synthetic_state_dict(synth.parseq_manifest(), 4321)
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(99)
    images = torch.randn((3, 3, 32, 128), generator=g)
    with torch.no_grad():
        ref_logits = model(images)
        ref_txt, _ = model.tokenizer.decode(ref_logits.softmax(-1))
        got = PR.forward(sd, images)
        memory = model.encode(images)
    got_txt, _ = PR.Tokenizer(cfg["charset_train"]).decode(got.softmax(-1))
    print("parseq logits", tuple(ref_logits.shape), "restated rel", rel(got, ref_logits), ref_txt, got_txt)
    assert got.shape == ref_logits.shape and rel(got, ref_logits) < 1e-5 and got_txt == ref_txt
    assert cfg["charset_train"] == PR.CHARSET_94
    tok_ref = model.tokenizer.encode(["Hello", "B200!"])
    assert torch.equal(tok_ref, PR.Tokenizer().encode(["Hello", "B200!"]))
    # crop preprocessing of ParseqPredictor.forward vs torchvision
    from torchvision import transforms
    tf = transforms.Compose([transforms.Resize((32, 128), transforms.InterpolationMode.BICUBIC, antialias=True),
                             transforms.Normalize(0.5, 0.5)])
    crops = [torch.rand((3, 57, 203), generator=g), torch.rand((3, 128, 384), generator=g)]
    ref_pre = torch.cat([tf(t[None]) for t in crops])
    assert rel(PR.preprocess(crops), ref_pre) < 1e-6
    with torch.no_grad():            # one image whose AR loop stops early (EOS at step 3): the refinement still queries 26 positions
        ref_single = model(images[2:3])
        got_single = PR.forward(sd, images[2:3])
    print("parseq early stop", tuple(ref_single.shape), rel(got_single, ref_single))
    assert got_single.shape == ref_single.shape and rel(got_single, ref_single) < 1e-5
    torch.save({"seed": 4321, "images": images.half(), "logits": ref_logits, "text": ref_txt, "memory_f16": memory.half(),
                "logits_single": ref_single,
                "crops_seed": 99, "pre": ref_pre.half()}, os.path.join(GOLD, "parseq.pt"))


if __name__ == "__main__":
    what = sys.argv[1:] or ["tiny"]
    if "parseq" in what:
        gen_parseq()
    os.makedirs(GOLD, exist_ok=True)
    if "tiny" in what:
        gen_tiny()
    if "loss" in what:
        gen_loss()
    if "noise_search" in what:
        gen_noise_search()
    m = sd = None
    if "full_unet" in what:
        m, sd = gen_full_unet()
    if "c1" in what:
        gen_c1(m, sd)
