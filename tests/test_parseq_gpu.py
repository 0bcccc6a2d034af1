"""PARSeq OCR scoring (reference sgm/modules/predictors/model.py:7-57, test.py:58-91) on the GPU: the
kernel-built recogniser (udifftext_b200/parseq.py behind the drop-in ParseqPredictor) against the fp32 restatement
(oracle/parseq_restated.py) and against the golden of the unmodified reference PARSeq (tests/golden/parseq.pt)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("b,lq,lk,heads,dh,masked", [(3, 26, 26, 12, 32, True), (2, 1, 7, 12, 32, True), (4, 26, 128, 12, 32, False),
                                                   (1, 5, 160, 2, 64, True)])
def test_mha_masked_matches_torch(udt_lib, b, lq, lk, heads, dh, masked):
    from udifftext_b200 import ops
    g = torch.Generator().manual_seed(lq * 1000 + lk)
    d = heads * dh
    q = torch.randn((b * lq, d), generator=g).cuda().half()
    kv = torch.randn((b * lk, 2 * d), generator=g).cuda().half()
    mask = kpm = None
    if masked:
        mask = torch.triu(torch.full((lq, lk), float("-inf")), 1).cuda()
        kpm = (torch.rand((b, lk), generator=g) < 0.2)
        kpm[:, 0] = False
        kpm = kpm.to(torch.uint8).cuda()
    got = ops.mha_masked(q, kv[:, :d], kv[:, d:], b, lq, lk, heads, mask=mask, kpm=kpm)
    split = lambda t, l: t.float().reshape(b, l, heads, dh).transpose(1, 2)
    s = split(q, lq) @ split(kv[:, :d], lk).transpose(-1, -2) / dh ** 0.5
    if masked:
        s = s + mask
        s = s.masked_fill(kpm.bool()[:, None, None, :], float("-inf"))
    ref = (s.softmax(-1) @ split(kv[:, d:], lk)).transpose(1, 2).reshape(b * lq, d)
    assert _rel(got, ref) < 2e-3


def test_gelu_epilogue(udt_lib):
    from udifftext_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn((300, 384), generator=g).cuda().half()
    w = (torch.randn((1536, 384), generator=g) / 384 ** 0.5).cuda().half()
    bias = torch.randn((1536,), generator=g).cuda()
    got = ops.linear(x, w, bias, act=ops.UDT_ACT_GELU)
    ref = torch.nn.functional.gelu(x.float() @ w.float().t() + bias)
    assert _rel(got, ref) < 1e-3
    got1 = ops.linear(x[:1], w, bias, act=ops.UDT_ACT_GELU)       # a single row (the AR decoder's M = batch)
    assert _rel(got1, ref[:1]) < 1e-3


def test_parseq_matches_oracle_and_reference_golden(udt_lib):
    from oracle import parseq_restated as PR
    from udifftext_b200 import synth
    from udifftext_b200.host.predictor import ParseqPredictor
    gold = torch.load(os.path.join(GOLD, "parseq.pt"))
    sd = synth.synthetic_state_dict(synth.parseq_manifest(), gold["seed"])
    pred = ParseqPredictor(state_dict=sd)
    pred.parseq = pred.parseq.to(torch.device("cuda", 0))          # test.py:60
    images = gold["images"].float().cuda()
    got = pred.parseq(images)
    mem = pred.parseq.exec.encode(images)
    torch.cuda.synchronize()
    e_mem = _rel(mem.view(3, 128, 384), gold["memory_f16"])
    e = _rel(got, gold["logits"])
    txt = pred.parseq.tokenizer.decode(got.softmax(-1))[0]
    print(f"PARSeq: encoder memory rel-L2 {e_mem:.3e}, logits rel-L2 {e:.3e}, text {txt} vs {gold['text']}")
    assert tuple(got.shape) == tuple(gold["logits"].shape)
    assert e_mem < 5e-3 and e < 2e-2
    # greedy strings: positions whose top-2 logit gap in the reference exceeds the fp16 error must decode identically
    ref = gold["logits"]
    top2 = ref.topk(2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) > 4 * (got.cpu() - ref).abs().max()
    assert torch.equal(got.cpu().argmax(-1)[clear], ref.argmax(-1)[clear])
    # the early-stopping single image (AR loop ends at its EOS, refinement still queries all 26 positions)
    single = pred.parseq(images[2:3])
    assert tuple(single.shape) == tuple(gold["logits_single"].shape) and _rel(single, gold["logits_single"]) < 2e-2
    # img2txt on variable-size crops: same preprocessing as torchvision's Resize + Normalize (golden `pre`)
    g = torch.Generator().manual_seed(gold["crops_seed"])
    torch.randn((3, 3, 32, 128), generator=g)
    crops = [torch.rand((3, 57, 203), generator=g), torch.rand((3, 128, 384), generator=g)]
    pre = torch.cat([pred.parseq_transform(t[None].cuda()) for t in crops])
    assert _rel(pre, gold["pre"].float()) < 1e-3
    out = pred.img2txt(crops)
    with torch.no_grad():
        ref_txt = PR.Tokenizer().decode(PR.forward({k: v.cuda() for k, v in sd.items()}, PR.preprocess([c.cuda() for c in crops])).softmax(-1))[0]
    print("img2txt:", out, "oracle:", ref_txt)
    assert len(out) == 2 and all(isinstance(s, str) for s in out)
    loss = pred.calc_loss(crops, ["Hello", "B200"])
    assert tuple(loss.shape) == (2,) and torch.isfinite(loss).all() and float(loss.max()) <= 1.0


def test_ocr_scoring_flow_of_test_py(udt_lib, tmp_path):
    """the OCR leg of test.py:58-91 on the device: predictor from `predictor_config` (configs/test.yaml:31-34) through the
    drop-in's instantiate_from_config, `.parseq.to(sampler.device)`, predict, crop by r_bbox, img2txt"""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "udifftext_b200", "dropin"))
    from sgm.util import instantiate_from_config
    from udifftext_b200 import api, synth
    ckpt = tmp_path / "parseq.pt"
    torch.save(synth.synthetic_state_dict(synth.parseq_manifest(), 4321), ckpt)
    cfgs = api.runtime_config(steps=3, batch_size=2, ocr_enabled=True,
                              predictor_config={"target": "sgm.modules.predictors.model.ParseqPredictor",
                                                "params": {"ckpt_path": str(ckpt)}})
    model = api.build_engine("tiny", torch.device("cuda", 0))
    sampler = api.init_sampling(cfgs)
    sampler.verbose = False
    predictor = instantiate_from_config(cfgs.predictor_config)
    predictor.parseq = predictor.parseq.to(sampler.device)
    batch = synth.synthetic_batch(9, 2, 64, 64, 5)
    batch["r_bbox"] = torch.tensor([[16, 32, 8, 56]] * 2)
    torch.manual_seed(3)
    results, _ = api.predict(cfgs, model, sampler, batch)
    crops = [results[i, :, t:b, l:r] for i, (t, b, l, r) in enumerate(batch["r_bbox"].tolist())]
    txt = predictor.img2txt(crops)
    assert len(txt) == 2 and all(isinstance(s, str) and len(s) <= 26 for s in txt)
    assert predictor.img2txt(crops) == txt            # deterministic
