"""Drop-in `sgm` package: put this directory (udifftext_b200/dropin) on sys.path AHEAD of the reference checkout and
the reference's own `util.py`, `test.py` and `demo.py` run unchanged on the B200 kernels.  Every module here only
re-exports the from-scratch implementations in `udifftext_b200.host` under the dotted paths the reference's YAML
configs and scripts use; components outside the inference hot path are not provided."""
from .util import instantiate_from_config  # noqa: F401
