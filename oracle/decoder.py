"""Oracle (test infrastructure): fairseq ``TransformerUnitDecoder`` restated in fp32 PyTorch (teacher forced).

The decoder is the immediate consumer of the hot path's output (reference call site
mm_s2ut/models/mm_s2s_transformer.py:693-696: ``self.decoder(prev_output_tokens, encoder_out=encoder_out)``).  It is
used on BOTH sides of the north-star "unit-argmax agreement >= 99 %" check: the same decoder weights are fed
the oracle's fp32 fused states and the CUDA path's fused states, and the arg-max units are compared.

Structure per fairseq ``s2ut_architecture_base``: embedding (V = target_code_size + 4
specials, padding_idx 1, scale sqrt(d)) + sinusoidal positions, 6 pre-LN layers (causal self-attention, encoder
attention with key padding mask, ReLU FFN), final LayerNorm, output projection tied to the embedding,
n_frames_per_step = 1.

Pinning: ``unit_decoder_forward`` is checked against HF ``Speech2TextDecoder`` (the port of the same fairseq decoder) with
copied weights, right-padded targets and an encoder padding mask (``make_golden.py: golden_hf_decoder`` ->
``tests/golden/hf_speech2text_decoder.npz``, ``test_unit_decoder_restatement_matches_hf_port``: 1e-7).  The criterion
below is restated from fairseq's source and unpinned.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from .s2t import make_positions, sinusoidal_table

Tensor = torch.Tensor


def init_decoder(d: int, ffn: int, layers: int, vocab: int = 1004, seed: int = 0) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)

    def xav(o, i, gain=1.0):
        a = gain * math.sqrt(6.0 / (i + o))
        return (torch.rand(o, i, generator=g) * 2 - 1) * a

    sd = {"embed_tokens.weight": torch.randn(vocab, d, generator=g) * d ** -0.5}
    sd["embed_tokens.weight"][1] = 0
    for i in range(layers):
        p = f"layers.{i}."
        for att in ("self_attn", "encoder_attn"):
            for proj, gain in (("q_proj", 2 ** -0.5), ("k_proj", 2 ** -0.5), ("v_proj", 2 ** -0.5), ("out_proj", 1.0)):
                sd[f"{p}{att}.{proj}.weight"] = xav(d, d, gain)
                sd[f"{p}{att}.{proj}.bias"] = torch.zeros(d)
        for ln in ("self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm"):
            sd[f"{p}{ln}.weight"], sd[f"{p}{ln}.bias"] = torch.ones(d), torch.zeros(d)
        sd[f"{p}fc1.weight"], sd[f"{p}fc1.bias"] = xav(ffn, d), torch.zeros(ffn)
        sd[f"{p}fc2.weight"], sd[f"{p}fc2.bias"] = xav(d, ffn), torch.zeros(d)
    sd["layer_norm.weight"], sd["layer_norm.bias"] = torch.ones(d), torch.zeros(d)
    return sd


def _mha(sd, p, q_in, kv_in, heads, key_padding_mask=None, causal=False, drop=lambda site, x: x, site=None):
    Tq, B, C = q_in.shape
    Tk = kv_in.shape[0]
    hd = C // heads
    q = F.linear(q_in, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"]) * hd ** -0.5
    k = F.linear(kv_in, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(kv_in, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    q = q.contiguous().view(Tq, B * heads, hd).transpose(0, 1)
    k = k.contiguous().view(Tk, B * heads, hd).transpose(0, 1)
    v = v.contiguous().view(Tk, B * heads, hd).transpose(0, 1)
    w = torch.bmm(q, k.transpose(1, 2))
    if causal:
        w = w + torch.triu(torch.full((Tq, Tk), float("-inf")), 1)[None]
    if key_padding_mask is not None:
        w = w.view(B, heads, Tq, Tk).masked_fill(key_padding_mask[:, None, None, :], float("-inf")).view(B * heads, Tq, Tk)
    a = torch.bmm(drop(site, F.softmax(w.float(), dim=-1)), v).transpose(0, 1).contiguous().view(Tq, B, C)
    return F.linear(a, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def unit_decoder_forward(sd: Dict[str, Tensor], prev_output_tokens: Tensor, encoder_out: Tensor,
                         encoder_padding_mask: Tensor, heads: int, drop=lambda site, x: x) -> Tensor:
    """prev_output_tokens [B, L] int64, encoder_out [T, B, d] -> logits [B, L, V].  ``drop(site, x)``: the training-mode
    dropout modules -- ("embed",), per layer i ("self", i), ("self_p", i), ("enc", i), ("enc_p", i), ("act", i), ("ffn", i);
    identity = eval mode."""
    d = encoder_out.shape[-1]
    B, L = prev_output_tokens.shape
    pos = make_positions(prev_output_tokens.ne(1), 1)
    # nn.Embedding(V, d, padding_idx=1): same values as plain indexing, no gradient for the padding row
    x = math.sqrt(d) * F.embedding(prev_output_tokens, sd["embed_tokens.weight"], padding_idx=1) + \
        sinusoidal_table(L + 2, d, 1)[pos]
    x = drop(("embed",), x.transpose(0, 1))
    i = 0
    while f"layers.{i}.fc1.weight" in sd:
        p = f"layers.{i}."
        r = x
        h = F.layer_norm(x, (d,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], 1e-5)
        x = r + drop(("self", i), _mha(sd, p + "self_attn.", h, h, heads, causal=True, drop=drop, site=("self_p", i)))
        r = x
        h = F.layer_norm(x, (d,), sd[p + "encoder_attn_layer_norm.weight"], sd[p + "encoder_attn_layer_norm.bias"], 1e-5)
        x = r + drop(("enc", i), _mha(sd, p + "encoder_attn.", h, encoder_out, heads,
                                      key_padding_mask=encoder_padding_mask, drop=drop, site=("enc_p", i)))
        r = x
        h = F.layer_norm(x, (d,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5)
        x = r + drop(("ffn", i), F.linear(drop(("act", i), F.relu(F.linear(h, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))),
                                          sd[p + "fc2.weight"], sd[p + "fc2.bias"]))
        i += 1
    x = F.layer_norm(x, (d,), sd["layer_norm.weight"], sd["layer_norm.bias"], 1e-5).transpose(0, 1)
    return F.linear(x, sd["embed_tokens.weight"])


def label_smoothed_nll_loss(logits: Tensor, target: Tensor, epsilon: float, padding_idx: int = 1):
    """fairseq ``label_smoothed_nll_loss`` (fairseq/criterions/label_smoothed_cross_entropy.py) on
    ``lprobs = log_softmax(logits.float())`` with ``ignore_index=padding_idx`` and ``reduce=True``: what the
    reference's criterion computes for the unit targets (mm_s2ut/criterions/speech_to_speech_criterion.py:58-72 via
    RdropLabelSmoothedCrossEntropyCriterion.compute_loss; scripts pass --label-smoothing 0.2).  Restated from the
    fairseq source as recalled (fairseq is not on this box): unpinned.  Returns (loss, nll_loss)."""
    lprobs = F.log_softmax(logits.float(), dim=-1).view(-1, logits.shape[-1])
    t = target.reshape(-1, 1)
    nll = -lprobs.gather(dim=-1, index=t)
    smooth = -lprobs.sum(dim=-1, keepdim=True)
    pad = t.eq(padding_idx)
    nll = nll.masked_fill(pad, 0.0).sum()
    smooth = smooth.masked_fill(pad, 0.0).sum()
    eps_i = epsilon / (lprobs.size(-1) - 1)
    return (1.0 - epsilon - eps_i) * nll + eps_i * smooth, nll
