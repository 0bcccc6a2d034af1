"""Pins the oracle (oracle/restated.py, the travelling fp32 restatement of the reference's algorithm) against the
golden vectors that oracle/make_golden.py produced by running the UNMODIFIED reference in the build container.
CPU only."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def tiny_sd():
    from udifftext_b200 import synth
    return synth.synthetic_state_dict(synth.load_manifest("tiny"), 1234)


def test_oracle_tiny_unet_and_attention_maps(tiny_sd):
    from oracle import restated as R
    from udifftext_b200 import synth
    gold = torch.load(os.path.join(GOLD, "tiny.pt"))
    probs = []
    with torch.no_grad():
        out = R.unet_forward(R._sub(tiny_sd, "model.diffusion_model."), gold["unet_x"], gold["unet_t"], gold["unet_ctx"],
                             model_channels=synth.ARCH["tiny"]["unet"]["model_channels"], probs_out=probs)
    assert _rel(out, gold["unet_out"]) < 1e-5
    assert out.abs().max() > 1e-2  # zero-initialised layers were re-randomised: parity is not vacuous
    assert len(probs) == len(gold["unet_probs"])
    for p, q in zip(probs, gold["unet_probs"]):
        assert (p - q).abs().max().item() < 1e-5


def test_oracle_tiny_predict_and_rng_order(tiny_sd):
    from oracle import restated as R
    from udifftext_b200 import synth
    gold = torch.load(os.path.join(GOLD, "tiny.pt"))
    batch = synth.synthetic_batch(gold["predict_config_id"], 2, 64, 64, None)
    torch.manual_seed(gold["predict_seed"])
    with torch.no_grad():
        img, z = R.predict(tiny_sd, batch, gold["predict_steps"], gold["predict_scale"])
    assert _rel(z, gold["predict_z"]) < 1e-4
    assert _rel(img, gold["predict_pixels"]) < 1e-4
    # conditioner: c / uc latents are two different posterior draws, in the reference's order
    torch.manual_seed(gold["predict_seed"])
    lat = (2, 4, 8, 8)
    n_c, n_uc = torch.randn(lat), torch.randn(lat)
    with torch.no_grad():
        c, uc = R.conditioner(R._sub(tiny_sd, "conditioner."), batch, n_c, n_uc)
    assert _rel(c["concat"], gold["c_concat"]) < 1e-5 and _rel(uc["concat"], gold["uc_concat"]) < 1e-5
    assert _rel(c["t_crossattn"], gold["c_crossattn"]) < 1e-5


def test_oracle_schedule_matches_reference_constants():
    """sigma schedule facts recorded from the reference run: 14.6146 ... 0.1345 ..., timesteps 999, 979, ..."""
    from oracle import restated as R
    sig = R.sampler_sigmas(50)
    assert sig.shape == (51,) and float(sig[-1]) == 0.0
    assert abs(float(sig[0]) - 14.6146) < 1e-3
    table = R.denoiser_sigmas()
    idx = R.sigma_to_idx(sig[:-1], table)
    assert idx.tolist() == list(range(999, 0, -20))
    assert torch.equal(table[idx], sig[:-1])  # sampler sigmas lie exactly on the denoiser's table


def test_oracle_full_unet_matches_reference_golden():
    """891.5 M-parameter UNet, batch 2, against the reference's CPU fp32 output (about half a minute on 8 cores)"""
    from oracle import restated as R
    from udifftext_b200 import synth
    gold = torch.load(os.path.join(GOLD, "full_unet.pt"))
    man = {k: v for k, v in synth.load_manifest("full").items() if k.startswith("model.diffusion_model.")}
    sd = R._sub(synth.synthetic_state_dict(man, 1234), "model.diffusion_model.")
    with torch.no_grad():
        out = R.unet_forward(sd, gold["x"], gold["t"], gold["ctx"])
    assert _rel(out, gold["out"]) < 1e-5


def test_oracle_min_local_loss_matches_reference_golden():
    """noise-search score (loss.py:192-235) against the reference's own FullLoss.get_min_local_loss output"""
    from oracle import restated as R
    gold = torch.load(os.path.join(GOLD, "loss.pt"))
    assert torch.allclose(R.gaussian_kernel(gold["kernel_size"], gold["sigma"]), gold["g_kernel"][0, 0], atol=1e-7)
    got = R.min_local_loss(gold["cache"], gold["mask"], gold["seg_mask"], gold["kernel_size"], gold["sigma"], gold["min_attn_size"])
    assert torch.allclose(got, gold["loss"], atol=1e-7)
    assert gold["loss"].abs().min() > 1e-3   # not vacuous: the masked region carries attention


def test_oracle_noise_search_matches_reference_golden(tiny_sd):
    """get_init_noise (sampling.py:264-322) restated vs the unmodified reference function's trial scores and winner"""
    from oracle import restated as R
    gold = torch.load(os.path.join(GOLD, "noise_search.pt"))
    from udifftext_b200 import synth
    batch = synth.synthetic_batch(gold["config_id"], 1, 64, 64, gold["label_len"])
    c = {"concat": gold["c_concat"], "t_crossattn": gold["c_crossattn"]}
    uc = {"concat": gold["uc_concat"], "t_crossattn": gold["uc_crossattn"]}
    with torch.no_grad():
        best, losses = R.init_noise_search(R._sub(tiny_sd, "model.diffusion_model."), gold["noises"], c, uc, batch["mask"],
                                           batch["seg_mask"], gold["scale"], gold["kernel_size"], gold["sigma"],
                                           gold["min_attn_size"])
    assert torch.allclose(losses, gold["losses"], atol=1e-5)
    assert torch.equal(best, gold["best"])
    assert (gold["losses"].max() - gold["losses"].min()).item() > 1e-3      # the trials are distinguishable


def test_parseq_restatement_matches_reference_golden():
    """oracle/parseq_restated.py vs the unmodified reference PARSeq run in the build container (tests/golden/parseq.pt:
    logits of 3 seeded images incl. one whose AR loop stops early, decoded strings, crop preprocessing vs torchvision)"""
    from oracle import parseq_restated as PR
    from udifftext_b200 import synth
    gold = torch.load(os.path.join(GOLD, "parseq.pt"))
    sd = synth.synthetic_state_dict(synth.parseq_manifest(), gold["seed"])
    images = gold["images"].float()
    with torch.no_grad():
        got = PR.forward(sd, images)
        single = PR.forward(sd, images[2:3])
    assert got.shape == gold["logits"].shape and _rel(got, gold["logits"]) < 2e-4      # inputs stored as fp16
    assert PR.Tokenizer().decode(got.softmax(-1))[0] == gold["text"]
    assert single.shape == gold["logits_single"].shape and _rel(single, gold["logits_single"]) < 2e-4
    g = torch.Generator().manual_seed(gold["crops_seed"])
    torch.randn((3, 3, 32, 128), generator=g)                                          # the images draw of make_golden
    crops = [torch.rand((3, 57, 203), generator=g), torch.rand((3, 128, 384), generator=g)]
    assert _rel(PR.preprocess(crops), gold["pre"].float()) < 1e-3
