import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # every fp32 oracle / torch reference evaluated on the GPU is TRUE fp32: TF32 has fp16's 10-bit mantissa, an oracle
    # computed with it carries an error of the size the parity tests measure
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="session")
def udt_lib():
    """Build (if stale) and load the C-ABI library."""
    from udifftext_b200 import build, lib

    build.build()
    return lib.load()
