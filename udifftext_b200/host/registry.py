"""Dotted-path registry: the reference's YAML / util.py name classes as `sgm.…`; they resolve to this package's
implementations without requiring the drop-in `sgm` package to shadow anything on sys.path."""
from __future__ import annotations

from . import autoencoder, conditioner, engine, loss, network, predictor, sampler, schedule

TARGETS = {
    "sgm.models.diffusion.DiffusionEngine": engine.DiffusionEngine,
    "sgm.models.autoencoder.AutoencoderKL": autoencoder.AutoencoderKL,
    "sgm.models.autoencoder.AutoencoderKLInferenceWrapper": autoencoder.AutoencoderKLInferenceWrapper,
    "sgm.modules.GeneralConditioner": conditioner.GeneralConditioner,
    "sgm.modules.encoders.modules.GeneralConditioner": conditioner.GeneralConditioner,
    "sgm.modules.encoders.modules.LabelEncoder": conditioner.LabelEncoder,
    "sgm.modules.encoders.modules.SpatialRescaler": conditioner.SpatialRescaler,
    "sgm.modules.encoders.modules.LatentEncoder": conditioner.LatentEncoder,
    "sgm.modules.diffusionmodules.openaimodel.UnifiedUNetModel": network.UnifiedUNetModel,
    "sgm.modules.diffusionmodules.wrappers.OpenAIWrapper": network.OpenAIWrapper,
    "sgm.modules.diffusionmodules.wrappers.IdentityWrapper": network.IdentityWrapper,
    "sgm.modules.diffusionmodules.denoiser.Denoiser": schedule.Denoiser,
    "sgm.modules.diffusionmodules.denoiser.DiscreteDenoiser": schedule.DiscreteDenoiser,
    "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling": schedule.EpsScaling,
    "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting": schedule.EpsWeighting,
    "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization": schedule.LegacyDDPMDiscretization,
    "sgm.modules.diffusionmodules.guiders.VanillaCFG": schedule.VanillaCFG,
    "sgm.modules.diffusionmodules.guiders.IdentityGuider": schedule.IdentityGuider,
    "sgm.modules.diffusionmodules.sampling_utils.NoDynamicThresholding": schedule.NoDynamicThresholding,
    "sgm.modules.diffusionmodules.sigma_sampling.DiscreteSampling": schedule.DiscreteSampling,
    "sgm.modules.diffusionmodules.sampling.EulerEDMSampler": sampler.EulerEDMSampler,
    "sgm.modules.diffusionmodules.loss.FullLoss": loss.FullLoss,
    "sgm.modules.predictors.model.ParseqPredictor": predictor.ParseqPredictor,
    "torch.nn.Identity": __import__("torch").nn.Identity,
}
