"""Oracle (test infrastructure): image fusion on top of the S2T encoder, fp32 PyTorch.

Restates, in eval-mode functional form over a fairseq-named ``state_dict``:

* ``SelectiveAttention.forward``     mm_s2ut/models/fuse.py:65-117 (num_heads == 1 as constructed at
                                      mm_s2ut/models/mm_s2s_transformer.py:129-140)
* ``MultimodalAttention.forward``    fuse.py:145-167 (nn.MultiheadAttention, kdim=vdim=image dim,
                                      num_heads=1, add_bias_kv=True, built at mm_s2s_transformer.py:141-155)
* ``fuse_img_feat``                   mm_s2s_transformer.py:594-622
* modality dropout + per-type sum     mm_s2s_transformer.py:496-530, :557-560

The golden vectors in tests/golden/fuse_*.npz were produced by the reference's own fuse.py
(oracle/make_golden.py) and pin these functions.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

from .s2t import s2t_encoder_forward

Tensor = torch.Tensor
_id = lambda t: t


def selective_attention(sd: Dict[str, Tensor], p: str, query: Tensor, key: Tensor, value: Tensor,
                        key_padding_mask: Optional[Tensor] = None, rnd: Callable = _id,
                        drop: Callable = lambda site, x: x):
    """query [Tq,B,d], key/value [Tk,B,Dk] -> (out [Tq,B,d], attn [B,Tq,Tk]); one head of width d."""
    Tq, B, _ = query.shape
    q = F.linear(rnd(query), rnd(sd[p + "q_proj.weight"]), sd[p + "q_proj.bias"])
    k = F.linear(rnd(key), rnd(sd[p + "k_proj.weight"]), sd[p + "k_proj.bias"])
    v = F.linear(rnd(value), rnd(sd[p + "v_proj.weight"]), sd[p + "v_proj.bias"])
    d = q.shape[-1]
    q = q * d ** -0.5                                   # fuse.py:80, qkhead_dim = attn_dim // 1
    q, k, v = q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1)
    attn = rnd(q) @ rnd(k).transpose(-2, -1)            # [B, Tq, Tk]
    if key_padding_mask is not None:
        attn = attn.masked_fill(key_padding_mask[:, None, :].to(torch.bool), float("-inf"))
    attn = attn.softmax(dim=-1)
    attn = drop(("sa_attn_p",), attn)                   # self.attn_drop (fuse.py:111), [B, Tq, Tk]
    x = (rnd(attn) @ rnd(v)).transpose(0, 1).contiguous()
    x = F.linear(rnd(x), rnd(sd[p + "proj.weight"]), sd[p + "proj.bias"])
    return x, attn


def multimodal_attention(sd: Dict[str, Tensor], p: str, text: Tensor, img: Tensor,
                         img_mask: Optional[Tensor] = None, rnd: Callable = _id,
                         drop: Callable = lambda site, x: x) -> Tensor:
    """nn.MultiheadAttention(embed_dim=d, num_heads=1, kdim=vdim=Dk, add_bias_kv=True), eval mode."""
    Tq, B, d = text.shape
    bq, bk, bv = sd[p + "in_proj_bias"].chunk(3)
    q = F.linear(rnd(text), rnd(sd[p + "q_proj_weight"]), bq)
    k = F.linear(rnd(img), rnd(sd[p + "k_proj_weight"]), bk)
    v = F.linear(rnd(img), rnd(sd[p + "v_proj_weight"]), bv)
    k = torch.cat([k, sd[p + "bias_k"].repeat(1, B, 1)], dim=0)      # one learned extra key/value
    v = torch.cat([v, sd[p + "bias_v"].repeat(1, B, 1)], dim=0)
    if img_mask is not None:
        img_mask = F.pad(img_mask.to(torch.bool), (0, 1))
    q = q * d ** -0.5
    q, k, v = q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1)
    attn = rnd(q) @ rnd(k).transpose(-2, -1)
    if img_mask is not None:
        attn = attn.masked_fill(img_mask[:, None, :], float("-inf"))
    attn = attn.softmax(dim=-1)
    attn = drop(("sa_attn_p",), attn)                   # nn.MultiheadAttention(dropout=SA_attention_dropout)
    x = (rnd(attn) @ rnd(v)).transpose(0, 1).contiguous()
    return F.linear(rnd(x), rnd(sd[p + "out_proj.weight"]), sd[p + "out_proj.bias"])


def fuse_img_feat(sd, prefix: str, mm_cfg, text: Tensor, idx: int, image: Tensor,
                  image_mask: Optional[Tensor], rnd: Callable = _id, drop: Callable = lambda site, x: x) -> Tensor:
    """text [T,B,d], image [Tk,B,Dk] -> fused [T,B,d].  ``drop(("image",), x)`` = SA_image_dropout (:596); identity in
    eval mode; ("text",) = SA_text_dropout, ("sa_attn_p",) = SA_attention_dropout on the probabilities."""
    if mm_cfg.image_pre_norm:
        image = F.layer_norm(image, (image.shape[-1],), sd[prefix + "image_pre_norm_module.weight"],
                             sd[prefix + "image_pre_norm_module.bias"], 1e-5)
    image = drop(("image",), image)
    text = drop(("text",), text)                        # SA_text_dropout (:597)
    kind = mm_cfg.multimodal_attention_type
    if kind == "selective_attention":
        out, _ = selective_attention(sd, f"{prefix}selective_attns.{idx}.", text, image, image, image_mask, rnd, drop)
    elif kind == "multimodal_attention":
        out = multimodal_attention(sd, f"{prefix}multimodal_attns.{idx}.", text, image, image_mask, rnd, drop)
    else:
        raise NotImplementedError(kind)
    if mm_cfg.use_selective_gate:
        merge = torch.cat([out, text], dim=-1)           # attention output first, speech second (:613)
        gate = torch.sigmoid(F.linear(rnd(merge), rnd(sd[f"{prefix}gate_denses.{idx}.weight"]),
                                      sd[f"{prefix}gate_denses.{idx}.bias"]))
        return (1 - gate) * text + gate * out
    return text + out


def mm_encoder_forward(sd, mm_cfg, src_tokens: Tensor, src_lengths: Tensor, imgs_list: List[Tensor],
                       img_masks_list: List[Optional[Tensor]], num_heads: int, prefix: str = "",
                       return_all_hiddens: bool = False, training: bool = False,
                       draws: Optional[tuple] = None, rnd: Callable = _id, drop: Callable = lambda site, x: x):
    """MM_S2STransformerEncoder.forward on the plain-S2T + fusion-at-top branch.

    ``training`` only switches the modality-dropout glue on (ordinary dropouts stay off so the result
    is deterministic); ``draws`` = (modality_drop_prob, audio_drop_prob), the two per-batch uniform
    draws the reference takes from ``np.random.random()`` (:497).  The reference's audio-drop branch
    raises NameError (:500); the evident intent -- zero the speech states -- is what is restated.
    A batch without padding makes the reference raise IndexError (:527); here it is an all-False mask.
    """
    out = s2t_encoder_forward(sd, src_tokens, src_lengths, num_heads, prefix, return_all_hiddens, rnd=rnd, drop=drop)
    if mm_cfg is None or not mm_cfg.is_fusion_top or not imgs_list:
        return out
    out["encoder_out"][0] = fusion_top(sd, prefix, mm_cfg, out["encoder_out"][0], imgs_list, img_masks_list, training,
                                       draws, rnd, drop)
    return out


def fusion_top(sd, prefix: str, mm_cfg, text: Tensor, imgs_list: List[Tensor], img_masks_list: List[Optional[Tensor]],
               training: bool = False, draws: Optional[tuple] = None, rnd: Callable = _id,
               drop: Callable = lambda site, x: x) -> Tensor:
    """The fusion-at-top statements of the reference's forward (mm_s2s_transformer.py:496-560): modality dropout,
    ``fuse_img_feat`` per image type, sum over the types.  text [T,B,d]; imgs_list entries [B,Tk,Dk].  Pinned by
    tests/golden/fuse_img_feat_*.npz (``glue_*`` keys: the reference's own statements, ast-extracted)."""
    imgs_list = list(imgs_list)
    if training and draws is not None:
        p_mod, p_aud = draws
        if p_mod < mm_cfg.modality_dropout:
            if p_aud < mm_cfg.audio_dropout:
                text = torch.zeros_like(text)
            else:
                imgs_list = [torch.zeros_like(i) for i in imgs_list]
    xs = []
    for idx, (img, img_mask) in enumerate(zip(imgs_list, img_masks_list)):
        xs.append(fuse_img_feat(sd, prefix, mm_cfg, text, idx, img.transpose(0, 1), img_mask, rnd, drop))
    res = xs[0]
    for x in xs[1:]:
        res = res + x
    return res
