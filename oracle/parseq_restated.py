"""fp32 PyTorch restatement of PARSeq inference — the OCR model behind `sgm.modules.predictors.model.ParseqPredictor`
(reference sgm/modules/predictors/model.py:7-57, used by test.py:58-91).  TEST INFRASTRUCTURE ONLY (same rules as
oracle/restated.py: imported by tests/, never by the product).

What is restated (file:line relative to /root/reference/src/parseq):
  * strhub/models/parseq/system.py:83-151  PARSeq.decode / PARSeq.forward (AR decoding + one cloze refinement iteration)
  * strhub/models/parseq/modules.py:27-108 DecoderLayer (two-stream pre-LN layer), Decoder; :125-133 TokenEmbedding
  * strhub/data/utils.py:46-127            Tokenizer (EOS = 0, charset, BOS, PAD), greedy decode + truncate at EOS
  * the ViT encoder: strhub/models/parseq/modules.py:111-122 subclasses `timm.models.vision_transformer.VisionTransformer`.
    timm is a THIRD-PARTY dependency that is not vendored in the reference and not installed here (requirements.txt:23 pins
    timm==0.9.2; src/parseq/requirements.txt:4 timm~=0.6.5), so its published algorithm is restated: patch embedding
    (Conv2d, kernel = stride = patch) -> + pos_embed (no class token) -> depth x pre-LN blocks [x += proj(MHA(LN(x)));
    x += fc2(GELU(fc1(LN(x))))] with LayerNorm eps 1e-6 and qkv bias -> final LayerNorm; all tokens returned
    (class_token=False, global_pool='', num_classes=0).

Pinning (oracle/make_golden.py `parseq`): the UNMODIFIED reference classes PARSeq / Decoder / DecoderLayer / TokenEmbedding /
Tokenizer are imported in the build container with `RestatedViT` below standing in for the absent timm class, run on
seeded weights and inputs, and their logits / decoded strings committed as tests/golden/parseq.pt; this file must
reproduce them.  The decoder, the AR / refinement loop and the tokenizer are therefore pinned against reference code;
the ViT arithmetic is "parity unpinned" (restated from timm's published definition only).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

CHARSET_94 = "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~"
# configs/main.yaml + configs/model/parseq.yaml + configs/charset/94_full.yaml (the `parseq` hub entry of hubconf.py:17-24)
HPARAMS = dict(img_size=(32, 128), patch_size=(4, 8), embed_dim=384, enc_num_heads=6, enc_mlp_ratio=4, enc_depth=12,
               dec_num_heads=12, dec_mlp_ratio=4, dec_depth=1, max_label_length=25, decode_ar=True, refine_iters=1)


class Tokenizer:
    """strhub/data/utils.py:46-127: ids = [E] (0), charset (1..), [B], [P]"""

    def __init__(self, charset: str = CHARSET_94):
        self.itos = ("[E]",) + tuple(charset) + ("[B]", "[P]")
        self.stoi = {s: i for i, s in enumerate(self.itos)}
        self.eos_id, self.bos_id, self.pad_id = self.stoi["[E]"], self.stoi["[B]"], self.stoi["[P]"]

    def __len__(self) -> int:
        return len(self.itos)

    def encode(self, labels: Sequence[str], device=None) -> torch.Tensor:
        rows = [torch.as_tensor([self.bos_id] + [self.stoi[c] for c in y] + [self.eos_id], dtype=torch.long, device=device)
                for y in labels]
        return nn.utils.rnn.pad_sequence(rows, batch_first=True, padding_value=self.pad_id)

    def decode(self, token_dists: torch.Tensor) -> Tuple[List[str], List[torch.Tensor]]:
        out_s, out_p = [], []
        for dist in token_dists:
            probs, ids = dist.max(-1)
            ids = ids.tolist()
            eos = ids.index(self.eos_id) if self.eos_id in ids else len(ids)
            out_s.append("".join(self.itos[i] for i in ids[:eos]))
            out_p.append(probs[: eos + 1])
        return out_s, out_p


# --------------------------------------------------------------------------------------------- functional restatement
def _ln(sd: SD, name: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _lin(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def vit_encode(sd: SD, img: torch.Tensor, heads: int = 6) -> torch.Tensor:
    """timm VisionTransformer.forward_features (class_token=False, global_pool=''); `sd` relative to `encoder.`"""
    w = sd["patch_embed.proj.weight"]
    x = F.conv2d(img, w, sd["patch_embed.proj.bias"], stride=w.shape[-2:]).flatten(2).transpose(1, 2)
    x = x + sd["pos_embed"]
    b, n, d = x.shape
    i = 0
    while f"blocks.{i}.norm1.weight" in sd:
        p = f"blocks.{i}."
        qkv = _lin(sd, p + "attn.qkv", _ln(sd, p + "norm1", x, 1e-6)).reshape(b, n, 3, heads, d // heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(b, n, d)
        x = x + _lin(sd, p + "attn.proj", a)
        x = x + _lin(sd, p + "mlp.fc2", F.gelu(_lin(sd, p + "mlp.fc1", _ln(sd, p + "norm2", x, 1e-6))))
        i += 1
    return _ln(sd, "norm", x, 1e-6)


def mha(sd: SD, p: str, q_in: torch.Tensor, kv_in: torch.Tensor, heads: int, attn_mask: Optional[torch.Tensor] = None,
        key_padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.MultiheadAttention(batch_first=True) forward (packed in_proj, additive float mask, boolean key padding mask)"""
    d = q_in.shape[-1]
    w, bias = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = F.linear(q_in, w[:d], bias[:d])
    k = F.linear(kv_in, w[d: 2 * d], bias[d: 2 * d])
    v = F.linear(kv_in, w[2 * d:], bias[2 * d:])
    b, lq, _ = q.shape
    lk = k.shape[1]
    split = lambda t, l: t.reshape(b, l, heads, d // heads).transpose(1, 2)
    s = split(q, lq) @ split(k, lk).transpose(-1, -2) / math.sqrt(d // heads)
    if attn_mask is not None:
        s = s + attn_mask
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    o = (s.softmax(-1) @ split(v, lk)).transpose(1, 2).reshape(b, lq, d)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def forward_stream(sd: SD, p: str, tgt, tgt_norm, tgt_kv, memory, tgt_mask, kpm, heads: int) -> torch.Tensor:
    """modules.py:57-75"""
    tgt = tgt + mha(sd, p + "self_attn.", tgt_norm, tgt_kv, heads, tgt_mask, kpm)
    tgt = tgt + mha(sd, p + "cross_attn.", _ln(sd, p + "norm1", tgt, 1e-5), memory, heads)
    return tgt + _lin(sd, p + "linear2", F.gelu(_lin(sd, p + "linear1", _ln(sd, p + "norm2", tgt, 1e-5))))


def decode(sd: SD, tgt: torch.Tensor, memory: torch.Tensor, tgt_mask=None, tgt_padding_mask=None, tgt_query=None,
           tgt_query_mask=None, heads: int = 12) -> torch.Tensor:
    """system.py:83-95 + Decoder.forward modules.py:99-108 (dropout off)"""
    n, l = tgt.shape
    d = memory.shape[-1]
    emb = lambda t: math.sqrt(d) * sd["text_embed.embedding.weight"][t]
    null_ctx = emb(tgt[:, :1])
    content = torch.cat([null_ctx, sd["pos_queries"][:, : l - 1] + emb(tgt[:, 1:])], dim=1)
    query = sd["pos_queries"][:, :l].expand(n, -1, -1) if tgt_query is None else tgt_query
    i = 0
    while f"decoder.layers.{i}.norm_q.weight" in sd:
        p = f"decoder.layers.{i}."
        last = f"decoder.layers.{i + 1}.norm_q.weight" not in sd
        qn, cn = _ln(sd, p + "norm_q", query, 1e-5), _ln(sd, p + "norm_c", content, 1e-5)
        query = forward_stream(sd, p, query, qn, cn, memory, tgt_query_mask, tgt_padding_mask, heads)
        if not last:
            content = forward_stream(sd, p, content, cn, cn, memory, tgt_mask, tgt_padding_mask, heads)
        i += 1
    return _ln(sd, "decoder.norm", query, 1e-5)


def forward(sd: SD, images: torch.Tensor, tok: Optional[Tokenizer] = None, max_label_length: int = 25,
            refine_iters: int = 1, enc_heads: int = 6, dec_heads: int = 12) -> torch.Tensor:
    """PARSeq.forward with decode_ar=True, max_length=None (testing): logits [B, <= 26, 95] (system.py:97-151)"""
    tok = tok or Tokenizer()
    dev = images.device
    bs = images.shape[0]
    num_steps = max_label_length + 1
    memory = vit_encode({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, images, enc_heads)
    pos_queries = sd["pos_queries"][:, :num_steps].expand(bs, -1, -1)
    tgt_mask = query_mask = torch.triu(torch.full((num_steps, num_steps), float("-inf"), device=dev), 1)
    tgt_in = torch.full((bs, num_steps), tok.pad_id, dtype=torch.long, device=dev)
    tgt_in[:, 0] = tok.bos_id
    logits = []
    for i in range(num_steps):
        j = i + 1
        out = decode(sd, tgt_in[:, :j], memory, tgt_mask[:j, :j], tgt_query=pos_queries[:, i:j],
                     tgt_query_mask=query_mask[i:j, :j], heads=dec_heads)
        p_i = _lin(sd, "head", out)
        logits.append(p_i)
        if j < num_steps:
            tgt_in[:, j] = p_i.squeeze(1).argmax(-1)
            if (tgt_in == tok.eos_id).any(dim=-1).all():
                break
    logits = torch.cat(logits, dim=1)
    if refine_iters:
        query_mask = query_mask.clone()
        query_mask[torch.triu(torch.ones(num_steps, num_steps, dtype=torch.bool, device=dev), 2)] = 0
        bos = torch.full((bs, 1), tok.bos_id, dtype=torch.long, device=dev)
        for _ in range(refine_iters):
            tgt_in = torch.cat([bos, logits[:, :-1].argmax(-1)], dim=1)
            kpm = (tgt_in == tok.eos_id).int().cumsum(-1) > 0
            l = tgt_in.shape[1]          # < num_steps when the AR loop stopped early; ALL num_steps positions are still queried
            out = decode(sd, tgt_in, memory, tgt_mask[:l, :l], kpm, tgt_query=pos_queries,
                         tgt_query_mask=query_mask[:, :l], heads=dec_heads)
            logits = _lin(sd, "head", out)
    return logits


def preprocess(crops: Sequence[torch.Tensor], img_size=(32, 128)) -> torch.Tensor:
    """ParseqPredictor.forward's transform (predictors/model.py:14-17,29): Resize(img_size, BICUBIC, antialias) + Normalize(.5,.5)
    per crop [3, h, w] in [0, 1] -> [B, 3, 32, 128]"""
    out = [F.interpolate(t[None].float(), size=tuple(img_size), mode="bicubic", antialias=True, align_corners=False) for t in crops]
    return (torch.cat(out) - 0.5) / 0.5


# --------------------------------------------------------------------------------------------- timm stand-in (pinning only)
class _Attn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        b, n, d = x.shape
        qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, d // self.num_heads).permute(2, 0, 3, 1, 4)
        return self.proj(F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(b, n, d))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim, heads, ratio):
        super().__init__()
        self.norm1, self.norm2 = nn.LayerNorm(dim, eps=1e-6), nn.LayerNorm(dim, eps=1e-6)
        self.attn, self.mlp = _Attn(dim, heads), _Mlp(dim, int(dim * ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.num_patches = (img_size[0] // patch_size[0]) * (img_size[1] // patch_size[1])
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=tuple(patch_size), stride=tuple(patch_size))

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class RestatedViT(nn.Module):
    """stands in for timm.models.vision_transformer.VisionTransformer when the reference's PARSeq is instantiated for
    pinning (same constructor keywords as modules.py:113-118 passes, same parameter names as timm)"""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 qkv_bias=True, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, embed_layer=PatchEmbed, num_classes=0,
                 global_pool="", class_token=False, **_):
        super().__init__()
        assert qkv_bias and not class_token and num_classes == 0 and global_pool == ""
        self.patch_embed = embed_layer(img_size, patch_size, in_chans, embed_dim)
        self.pos_embed = nn.Parameter(torch.randn(1, self.patch_embed.num_patches, embed_dim) * 0.02)
        self.blocks = nn.Sequential(*[_Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)

    def no_weight_decay(self):
        return {"pos_embed"}

    def forward_features(self, x):
        return self.norm(self.blocks(self.patch_embed(x) + self.pos_embed))
