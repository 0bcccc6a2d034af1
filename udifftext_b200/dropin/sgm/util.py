from udifftext_b200.host.config import default, get_obj_from_str, instantiate_from_config  # noqa: F401
from udifftext_b200.host.schedule import append_dims, append_zero  # noqa: F401


def disabled_train(self, mode=True):
    """kept for API compatibility: the engine is inference-only"""
    return self
