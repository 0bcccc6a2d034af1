from udifftext_b200.host.network import IdentityWrapper, OpenAIWrapper  # noqa: F401
OPENAIUNETWRAPPER = "sgm.modules.diffusionmodules.wrappers.OpenAIWrapper"
