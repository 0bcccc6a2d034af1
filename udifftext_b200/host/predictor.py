"""ParseqPredictor — OCR scoring of generated crops (reference: sgm/modules/predictors/model.py:7-57; instantiated by
test.py:58-60 from configs/test.yaml:31-34 and by FullLoss when `ocr_enabled`).  Same constructor, `forward`, `img2txt`,
`calc_loss`, `freeze`, `.parseq` (with `.tokenizer`, `.hparams.img_size`, `.to(device)`) as the reference; the network
runs on the sm_100a kernels (udifftext_b200/parseq.py).

The reference builds the model with `torch.hub.load('./src/parseq', 'parseq', source='local')` and then loads
`ckpt_path` (a plain state_dict, `parseq-bb5792a6.pt`).  Here the state_dict alone defines the model.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .config import AttrDict

CHARSET_94 = "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~"


class Tokenizer:
    """src/parseq/strhub/data/utils.py:46-127: ids = [E] (0), the charset (1..), [B], [P]; greedy decode, cut at [E]"""

    BOS, EOS, PAD = "[B]", "[E]", "[P]"

    def __init__(self, charset: str = CHARSET_94):
        self._itos = (self.EOS,) + tuple(charset) + (self.BOS, self.PAD)
        self._stoi = {s: i for i, s in enumerate(self._itos)}
        self.eos_id, self.bos_id, self.pad_id = (self._stoi[s] for s in (self.EOS, self.BOS, self.PAD))

    def __len__(self) -> int:
        return len(self._itos)

    def encode(self, labels: Sequence[str], device=None) -> torch.Tensor:
        rows = [[self.bos_id] + [self._stoi[c] for c in y] + [self.eos_id] for y in labels]
        width = max(len(r) for r in rows)
        return torch.tensor([r + [self.pad_id] * (width - len(r)) for r in rows], dtype=torch.long, device=device)

    def decode(self, token_dists: torch.Tensor, raw: bool = False) -> Tuple[List, List[torch.Tensor]]:
        batch_tokens, batch_probs = [], []
        for dist in token_dists:
            probs, ids = dist.max(-1)
            ids = ids.tolist()
            if not raw:
                eos = ids.index(self.eos_id) if self.eos_id in ids else len(ids)
                ids, probs = ids[:eos], probs[: eos + 1]
            toks = [self._itos[i] for i in ids]
            batch_tokens.append(toks if raw else "".join(toks))
            batch_probs.append(probs)
        return batch_tokens, batch_probs


class _Parseq:
    """what `predictor.parseq` has to offer the callers (test.py:60, predictors/model.py:14,27,35,43)"""

    def __init__(self, sd: Dict[str, torch.Tensor]):
        self._sd = {k: v.detach().cpu() for k, v in sd.items()}
        w = self._sd["encoder.patch_embed.proj.weight"]
        n = self._sd["encoder.pos_embed"].shape[1]
        ph, pw = int(w.shape[2]), int(w.shape[3])
        gh = max(1, int(round((n * pw / (4 * ph)) ** 0.5)))          # 1:4 aspect (32 x 128 for the shipped model)
        self.hparams = AttrDict(img_size=[gh * ph, (n // gh) * pw])
        ntok = self._sd["text_embed.embedding.weight"].shape[0]
        if ntok != len(CHARSET_94) + 3:
            raise ValueError(f"PARSeq checkpoint with {ntok} tokens: only the 94-character charset is supported")
        self.tokenizer = Tokenizer(CHARSET_94)
        self.exec = None
        self.device: Optional[torch.device] = None

    def to(self, device):
        from ..parseq import ParseqB200
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("PARSeq runs on the sm_100a kernels only (no CPU path)")
        if self.exec is None or self.device != device:
            self.exec = ParseqB200(self._sd, device, max_label_length=self._sd["pos_queries"].shape[1] - 1)
            self.device = device
        return self

    def eval(self):
        return self

    def parameters(self):
        return iter(self._sd.values())

    def state_dict(self):
        return dict(self._sd)

    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        if self.exec is None:
            raise RuntimeError("PARSeq: call .to(cuda device) first (test.py:60 does)")
        return self.exec(images)


class ParseqPredictor:
    def __init__(self, ckpt_path=None, freeze=True, state_dict: Optional[Dict[str, torch.Tensor]] = None, *args, **kwargs):
        if state_dict is None:
            if ckpt_path is None:
                raise ValueError("ParseqPredictor needs ckpt_path (the PARSeq state_dict, e.g. parseq-bb5792a6.pt)")
            state_dict = torch.load(ckpt_path, map_location="cpu")
            state_dict = state_dict.get("state_dict", state_dict)
        self.parseq = _Parseq(state_dict)
        if freeze:
            self.freeze()

    def freeze(self):
        return self

    def to(self, device):
        self.parseq = self.parseq.to(device)
        return self

    def parameters(self):
        return self.parseq.parameters()

    def parseq_transform(self, t: torch.Tensor) -> torch.Tensor:
        """predictors/model.py:14-17: Resize(img_size, BICUBIC, antialias=True) + Normalize(0.5, 0.5) of one [1, 3, h, w] crop"""
        size = tuple(int(s) for s in self.parseq.hparams.img_size)
        y = F.interpolate(t.float(), size=size, mode="bicubic", antialias=True, align_corners=False)
        return (y - 0.5) / 0.5

    def forward(self, x) -> torch.Tensor:
        """x: sequence of crops [3, h, w] in [0, 1] (any sizes) -> logits [B, <= 26, 95] (predictors/model.py:27-32)"""
        dev = self.parseq.device
        if dev is None:
            raise RuntimeError("ParseqPredictor: move `.parseq` to the CUDA device first (test.py:60)")
        imgs = torch.cat([self.parseq_transform(t[None].to(dev)) for t in x])
        return self.parseq(imgs)

    __call__ = forward

    def img2txt(self, x) -> List[str]:
        pred = self(x)
        label, _ = self.parseq.tokenizer.decode(pred)
        return label

    def calc_loss(self, x, label) -> torch.Tensor:
        """predictors/model.py:41-57: per-sample cross-entropy of the predicted characters against `label`, clamped at 1"""
        preds = self(x)
        gt_ids = self.parseq.tokenizer.encode(label).to(preds.device)
        losses = []
        for pred, gt_id in zip(preds, gt_ids):
            eos_id = int((gt_id == 0).nonzero()[0].item())
            gt = gt_id[1:eos_id]
            pr = pred[: eos_id - 1, :]
            ce = F.cross_entropy(pr.permute(1, 0)[None], gt[None])
            losses.append(torch.clamp(ce, max=1.0)[None])
        return torch.cat(losses)
