"""Host-side mirror of the reference's plugin interface for the inference hot path: the classes the reference's
YAML configs and `util.py` / `test.py` / `demo.py` instantiate (`DiffusionEngine`, `EulerEDMSampler`,
`GeneralConditioner`, ...), with the same constructor arguments, methods, `state_dict` keys and RNG draw order,
written from scratch on top of the sm_100a kernels.  `udifftext_b200/dropin/sgm` re-exports them under the
reference's dotted paths."""
