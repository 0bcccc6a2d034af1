"""Helper of tests/test_host_logic.py::test_reference_test_py_and_demo_py_run_on_the_dropin (run in a subprocess, build
container only): executes the UNMODIFIED /root/reference/test.py and demo.py sources against the drop-in `sgm` package.

The reference's entry points hard-code CUDA devices and need checkpoints / datasets, none of which exist in the build
container, so the *device layer* is stubbed while every call the reference makes stays checked:
  * `model` / `sampler` are `unittest.mock.create_autospec` images of OUR real DiffusionEngine / GeneralConditioner /
    EulerEDMSampler instances — a call with a keyword our classes do not accept raises TypeError;
  * `torch.device("cuda", ...)` inside the reference's util.py resolves to the CPU;
  * third-party modules that are not installed (omegaconf, pytorch_lightning, gradio, lpips) and the reference's own dataset
    / metrics modules (out of scope) are stand-ins;
  * `configs/test.yaml` is the reference's shipped file (`ocr_enabled: True`): its `predictor_config` is instantiated through
    our `sgm.util.instantiate_from_config` into the real ParseqPredictor from a synthetic PARSeq state_dict.
Prints "OK <what>" lines; any exception fails the test.
"""
import os
import runpy
import sys
import tempfile
import types
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("UDT_REFERENCE_ROOT", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "udifftext_b200", "dropin"), ROOT, REF]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from udifftext_b200 import api, synth  # noqa: E402
from udifftext_b200.host.config import to_attr  # noqa: E402


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _yaml(path):
    import yaml
    with open(path) as f:
        return to_attr(yaml.safe_load(f))


class _OmegaConf:
    load = staticmethod(_yaml)
    create = staticmethod(to_attr)


stub("omegaconf", OmegaConf=_OmegaConf, ListConfig=list, DictConfig=dict)
stub("pytorch_lightning", seed_everything=lambda s: torch.manual_seed(s))
stub("lpips")
stub("metrics", calc_fid=lambda *a, **k: None, calc_lpips=lambda *a, **k: None)
stub("dataset")
stub("dataset.dataloader", get_dataloader=lambda *a, **k: [])


class _Blocks:
    def queue(self):
        return self


stub("gradio", Blocks=_Blocks)

# ---------------------------------------------------------------- our real objects -> autospec images (signature-checked)
H = W = 64
B = 2
engine = api.build_engine("tiny")                      # real DiffusionEngine (weights on the host; never executed here)
cfg_rt = api.runtime_config(steps=4, batch_size=B)
real_sampler = api.init_sampling(cfg_rt)

lat = (B, 4, H // 8, W // 8)
cond = {"t_crossattn": torch.zeros(B, 12, 128), "concat": torch.zeros(B, 5, H // 8, W // 8)}
model = mock.create_autospec(engine, instance=True)
model.conditioner = mock.create_autospec(engine.conditioner, instance=True)
model.conditioner.get_unconditional_conditioning.return_value = (cond, dict(cond))
model.decode_first_stage.return_value = torch.rand(B, 3, H, W) * 2 - 1


def make_sampler(*_a, **_k):
    s = mock.create_autospec(real_sampler, instance=True)
    s.device = torch.device("cpu")
    s.get_init_noise.return_value = torch.randn(lat)
    s.return_value = torch.randn(lat)                  # __call__
    return s


class _TorchCpu:
    """`torch` as util.py sees it: every device request lands on the CPU"""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*a, **k):
        return torch.device("cpu")


def load(path):
    cwd = os.getcwd()
    os.chdir(REF)                                       # ./configs/*.yaml of the reference
    try:
        ns = runpy.run_path(path, run_name="reference_entry")
    finally:
        os.chdir(cwd)
    for fn_name in ("prepare_batch", "init_sampling", "init_model"):
        ns[fn_name].__globals__["torch"] = _TorchCpu()  # util.py's module globals
    return ns


# ---------------------------------------------------------------- test.py: predict (19-40) and test (43-126)
ns = load(os.path.join(REF, "test.py"))
from sgm.modules.diffusionmodules.sampling import EulerEDMSampler  # noqa: E402  (the drop-in's)
assert ns["EulerEDMSampler"] is EulerEDMSampler and EulerEDMSampler.__module__ == "udifftext_b200.host.sampler"
sampler = ns["init_sampling"](cfg_rt)                  # the reference's factory builds OUR sampler
assert type(sampler) is EulerEDMSampler and sampler.guider.scale == 5.0
batch = synth.synthetic_batch(3, B, H, W, 5)
batch["r_bbox"] = torch.tensor([[16, 32, 8, 56]] * B)
samples, z = ns["predict"](cfg_rt, model, make_sampler(), dict(batch))
assert tuple(samples.shape) == (B, 3, H, W) and float(samples.min()) >= 0 and float(samples.max()) <= 1
kw = model.conditioner.get_unconditional_conditioning.call_args.kwargs
assert set(kw) == {"batch_uc", "force_uc_zero_embeddings"} and kw["batch_uc"]["label"] == [""] * B
print("OK test.py predict")

cfgs = _yaml(os.path.join(REF, "configs", "test.yaml"))       # the shipped file: ocr_enabled True
assert cfgs.ocr_enabled and cfgs.predictor_config.target == "sgm.modules.predictors.model.ParseqPredictor"
tmp = tempfile.mkdtemp()
ckpt = os.path.join(tmp, "parseq.pt")
torch.save(synth.synthetic_state_dict(synth.parseq_manifest(), 4321), ckpt)
cfgs.predictor_config.params.ckpt_path = ckpt
cfgs.output_dir, cfgs.temp_dir = os.path.join(tmp, "outputs"), os.path.join(tmp, "temp")
cfgs.batch_size, cfgs.steps, cfgs.noise_iters, cfgs.max_iter = B, 4, 0, 2
from udifftext_b200.host import predictor as _pred  # noqa: E402
seen = {}


def fake_to(self, device):                              # the kernels cannot run here: record the request, stay on the host
    seen["device"] = device
    return self


def fake_img2txt(self, crops):
    seen["crops"] = [tuple(c.shape) for c in crops]
    return [lab for lab in batch["label"]]


with mock.patch.object(_pred._Parseq, "to", fake_to), mock.patch.object(_pred.ParseqPredictor, "img2txt", fake_img2txt):
    ns["test"](model, make_sampler(), [dict(batch), dict(batch)], cfgs)
assert seen["crops"] == [(3, 16, 48)] * B and str(seen["device"]) == "cpu"
assert sorted(os.listdir(cfgs.output_dir)) == ["0.png", "fake", "real"]
print("OK test.py test() with configs/test.yaml (ocr_enabled)")

# ---------------------------------------------------------------- demo.py: predict (15-36) and demo_predict (39-116)
nd = load(os.path.join(REF, "demo.py"))
demo_cfg = _yaml(os.path.join(REF, "configs", "demo.yaml"))
demo_cfg.H = demo_cfg.W = H
g = nd["demo_predict"].__globals__
g.update(cfgs=demo_cfg, model=model, global_index=0, init_sampling=make_sampler)
model.decode_first_stage.return_value = torch.rand(3, 3, H, W) * 2 - 1
lat = (3, 4, H // 8, W // 8)
blk = {"image": np.zeros((96, 80, 3), np.uint8), "mask": np.zeros((96, 80, 3), np.uint8)}
blk["mask"][20:40, 10:70] = 255
results, attn_map, seg = nd["demo_predict"](blk, "Hello", 3, 4, 4.0, 7, False)
assert len(results) == 3 and results[0].size == (W, H) and attn_map is None and seg is None
assert demo_cfg.batch_size == 3 and demo_cfg.noise_iters == 0 and demo_cfg.scale[0] == 4.0
print("OK demo.py demo_predict")
