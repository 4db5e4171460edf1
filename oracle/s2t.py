"""Oracle (test infrastructure): fairseq S2T transformer encoder restated in fp32 PyTorch.

fairseq is not vendored by the reference and not installed here; this follows the upstream
``main`` semantics summarised in SURVEY.md §8-appendix:
``fairseq/models/speech_to_text/modules/convolution.py`` (Conv1dSubsampler),
``fairseq/models/speech_to_text/s2t_transformer.py`` (S2TTransformerEncoder._forward),
``fairseq/modules/{transformer_layer,multihead_attention,sinusoidal_positional_embedding}.py``.
The reference reaches it through ``super().forward`` at mm_s2ut/models/mm_s2s_transformer.py:464.

All functions take a plain ``state_dict`` with fairseq key names (``prefix`` = "encoder." or "").
``rnd`` is an optional operand-rounding hook (identity by default) used only by the precision
study in tests: it is applied to both operands of every contraction to emulate bf16 tensor-core
inputs with fp32 accumulation.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
_id = lambda t: t


def out_seq_lens(lens: Tensor, n_layers: int = 2) -> Tensor:
    out = lens.clone()
    for _ in range(n_layers):
        out = ((out.float() - 1) / 2 + 1).floor().long()
    return out


def lengths_to_padding_mask(lens: Tensor, max_len: Optional[int] = None) -> Tensor:
    m = int(lens.max()) if max_len is None else max_len
    return torch.arange(m)[None, :] >= lens[:, None]


def sinusoidal_table(num_embeddings: int, dim: int, padding_idx: int = 1) -> Tensor:
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    e = torch.arange(num_embeddings, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1).view(num_embeddings, -1)
    if dim % 2 == 1:
        e = torch.cat([e, torch.zeros(num_embeddings, 1)], dim=1)
    e[padding_idx, :] = 0
    return e


def make_positions(tokens_ne_pad: Tensor, padding_idx: int = 1) -> Tensor:
    mask = tokens_ne_pad.int()
    return (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + padding_idx


def conv1d_subsampler(sd: Dict[str, Tensor], prefix: str, src_tokens: Tensor, src_lengths: Tensor,
                      rnd: Callable = _id):
    """[B, T, 80] -> [T', B, d]: Conv1d(k=5, s=2, p=2) + GLU, twice."""
    x = src_tokens.transpose(1, 2).contiguous()
    i = 0
    while f"{prefix}subsample.conv_layers.{i}.weight" in sd:
        w = sd[f"{prefix}subsample.conv_layers.{i}.weight"]
        b = sd[f"{prefix}subsample.conv_layers.{i}.bias"]
        x = F.conv1d(rnd(x), rnd(w), b, stride=2, padding=w.shape[2] // 2)
        x = F.glu(x, dim=1)
        i += 1
    x = x.transpose(1, 2).transpose(0, 1).contiguous()
    return x, out_seq_lens(src_lengths, i)


def multihead_self_attention(sd, p: str, x: Tensor, key_padding_mask: Tensor, num_heads: int,
                             rnd: Callable = _id, drop: Callable = lambda site, x: x, index: int = 0) -> Tensor:
    T, B, C = x.shape
    hd = C // num_heads
    xq = rnd(x)
    q = F.linear(xq, rnd(sd[p + "q_proj.weight"]), sd[p + "q_proj.bias"]) * hd ** -0.5
    k = F.linear(xq, rnd(sd[p + "k_proj.weight"]), sd[p + "k_proj.bias"])
    v = F.linear(xq, rnd(sd[p + "v_proj.weight"]), sd[p + "v_proj.bias"])
    q = q.contiguous().view(T, B * num_heads, hd).transpose(0, 1)
    k = k.contiguous().view(T, B * num_heads, hd).transpose(0, 1)
    v = v.contiguous().view(T, B * num_heads, hd).transpose(0, 1)
    w = torch.bmm(rnd(q), rnd(k).transpose(1, 2)).view(B, num_heads, T, T)
    w = w.masked_fill(key_padding_mask[:, None, None, :], float("-inf")).view(B * num_heads, T, T)
    w = F.softmax(w.float(), dim=-1)
    w = drop(("attn_p", index), w)                 # MultiheadAttention.dropout_module on the probabilities [B*H, T, T]
    a = torch.bmm(rnd(w), rnd(v)).transpose(0, 1).contiguous().view(T, B, C)
    return F.linear(rnd(a), rnd(sd[p + "out_proj.weight"]), sd[p + "out_proj.bias"])


def _no_drop(site, x):
    return x


def encoder_layer(sd, p: str, x: Tensor, mask: Tensor, num_heads: int, rnd: Callable = _id,
                  drop: Callable = _no_drop, index: int = 0) -> Tensor:
    """Pre-LN TransformerEncoderLayer, ReLU.  ``drop(site, x)`` stands for the layer's three FairseqDropout modules
    (sites ("attn", i), ("act", i), ("ffn", i): after self-attention, after the activation, after fc2); identity in eval
    mode; ("attn_p", i) is the attention-probability dropout inside the MHA."""
    C = x.shape[-1]
    r = x
    h = F.layer_norm(x, (C,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], 1e-5)
    h = multihead_self_attention(sd, p + "self_attn.", h, mask, num_heads, rnd, drop, index)
    x = r + drop(("attn", index), h)
    r = x
    h = F.layer_norm(x, (C,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5)
    h = drop(("act", index), F.relu(F.linear(rnd(h), rnd(sd[p + "fc1.weight"]), sd[p + "fc1.bias"])))
    h = F.linear(rnd(h), rnd(sd[p + "fc2.weight"]), sd[p + "fc2.bias"])
    return r + drop(("ffn", index), h)


def s2t_encoder_forward(sd: Dict[str, Tensor], src_tokens: Tensor, src_lengths: Tensor, num_heads: int,
                        prefix: str = "", return_all_hiddens: bool = False, no_scale_embedding: bool = False,
                        rnd: Callable = _id, drop: Callable = _no_drop) -> Dict[str, List[Tensor]]:
    """S2TTransformerEncoder._forward.  src_tokens [B, T, 80] (post-CMVN, zero-padded).  ``drop(site, x)``: the
    training-mode dropout modules (site ("embed",) after the positions, then the layers'); identity = eval mode."""
    x, in_lens = conv1d_subsampler(sd, prefix, src_tokens, src_lengths, rnd)
    T, B, C = x.shape
    x = (1.0 if no_scale_embedding else math.sqrt(C)) * x
    mask = lengths_to_padding_mask(in_lens, T)
    pos_idx = make_positions(~mask, 1)                       # padded -> 1 (zero row), valid t -> t + 2
    table = sinusoidal_table(T + 2, C, 1)
    x = x + table[pos_idx.reshape(-1)].view(B, T, C).transpose(0, 1)
    x = drop(("embed",), x)
    states = []
    i = 0
    while f"{prefix}transformer_layers.{i}.fc1.weight" in sd:
        x = encoder_layer(sd, f"{prefix}transformer_layers.{i}.", x, mask, num_heads, rnd, drop, i)
        if return_all_hiddens:
            states.append(x)
        i += 1
    if f"{prefix}layer_norm.weight" in sd:
        x = F.layer_norm(x, (C,), sd[f"{prefix}layer_norm.weight"], sd[f"{prefix}layer_norm.bias"], 1e-5)
    return {
        "encoder_out": [x],
        "encoder_padding_mask": [mask] if mask.any() else [],
        "encoder_embedding": [],
        "encoder_states": states,
        "src_tokens": [],
        "src_lengths": [],
    }
