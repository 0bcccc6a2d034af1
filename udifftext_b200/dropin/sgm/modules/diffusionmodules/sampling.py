"""`from sgm.modules.diffusionmodules.sampling import *` (reference util.py:4) must yield EulerEDMSampler"""
from udifftext_b200.host.sampler import BaseDiffusionSampler, EDMSampler, EulerEDMSampler, SingleStepDiffusionSampler  # noqa: F401

__all__ = ["BaseDiffusionSampler", "SingleStepDiffusionSampler", "EDMSampler", "EulerEDMSampler"]
