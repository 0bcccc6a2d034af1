from udifftext_b200.host.conditioner import GeneralConditioner  # noqa: F401

UNCONDITIONAL_CONFIG = {"target": "sgm.modules.GeneralConditioner", "params": {"emb_models": []}}
