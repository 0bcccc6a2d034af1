"""Attention-map figure of `UnifiedUNetModel.save_attn_map` (openaimodel.py:575-589): a 3 x 4 grid of per-token heat maps
with the token as the title, written to `temp/attn_map/attn_map_<name>.png`, which `demo.py:104` opens.

The reference draws it with seaborn + matplotlib; neither is a dependency of this package, so the figure is rendered
with PIL: every map is scaled to its own [min, max] (seaborn's default vmin / vmax) and coloured with a dark-to-light
sequential ramp.  Presentation only — the numbers `demo.py` consumes come from `save_segment_map`'s .npy file.
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np

# anchors of a sequential "rocket"-like ramp (dark purple -> red -> cream), interpolated linearly
_RAMP = np.array([[3, 5, 26], [53, 20, 68], [120, 28, 109], [190, 40, 90], [234, 81, 63], [245, 150, 105],
                  [250, 215, 185], [250, 235, 221]], dtype=np.float32)


def _colour(m: np.ndarray) -> np.ndarray:
    lo, hi = float(m.min()), float(m.max())
    t = (m - lo) / (hi - lo) if hi > lo else np.zeros_like(m)
    pos = t * (len(_RAMP) - 1)
    i0 = np.clip(np.floor(pos).astype(np.int64), 0, len(_RAMP) - 2)
    f = (pos - i0)[..., None]
    return (_RAMP[i0] * (1 - f) + _RAMP[i0 + 1] * f).astype(np.uint8)


def save_attn_figure(attn_map: np.ndarray, tokens: Sequence[str], path: str, cell: int = 256, pad: int = 28) -> None:
    """attn_map [L, h, w] -> PNG with up to 12 panels (3 rows x 4 columns), panel j titled tokens[j]"""
    from PIL import Image, ImageDraw

    n = min(12, attn_map.shape[0])
    fig = Image.new("RGB", (4 * (cell + pad) + pad, 3 * (cell + pad) + pad), (255, 255, 255))
    draw = ImageDraw.Draw(fig)
    for j in range(n):
        r, c = divmod(j, 4)
        x0, y0 = pad + c * (cell + pad), pad + r * (cell + pad)
        panel = Image.fromarray(_colour(np.asarray(attn_map[j], dtype=np.float32))).resize((cell, cell), Image.NEAREST)
        fig.paste(panel, (x0, y0))
        if j < len(tokens):
            draw.text((x0 + cell // 2 - 4, y0 - 16), str(tokens[j]), fill=(0, 0, 0))
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    fig.save(path)
