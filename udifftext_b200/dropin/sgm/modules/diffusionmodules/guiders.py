from udifftext_b200.host.schedule import IdentityGuider, VanillaCFG  # noqa: F401
