"""Reference-order noise draws.  Every random tensor on the hot path is drawn on the CPU default generator and then
moved to the device (distributions.py:36-41, sampling.py:269): results depend only on the seed and on the ORDER and
SHAPES of the draws.  When a request batch is sharded over ranks, each rank draws the noise of the WHOLE batch and
keeps its own rows, so the images do not depend on the number of GPUs."""
from __future__ import annotations

import contextlib
from typing import Optional, Tuple

import torch

_shard: Optional[Tuple[int, int, int]] = None  # (global_batch, lo, hi)


@contextlib.contextmanager
def batch_shard(global_batch: int, lo: int, hi: int):
    """inside this context `randn((hi-lo, ...))` draws `(global_batch, ...)` and returns rows [lo, hi)"""
    global _shard
    prev = _shard
    _shard = (int(global_batch), int(lo), int(hi))
    try:
        yield
    finally:
        _shard = prev


def randn(shape, device) -> torch.Tensor:
    shape = tuple(int(s) for s in shape)
    if _shard is None:
        return torch.randn(shape).to(device)
    gb, lo, hi = _shard
    assert shape[0] == hi - lo, f"sharded draw: local batch {shape[0]} != shard size {hi - lo}"
    return torch.randn((gb,) + shape[1:])[lo:hi].contiguous().to(device)
