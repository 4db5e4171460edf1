"""``MM_S2UTTransformerModel`` without fairseq: encoder + unit decoder, same call surface as the reference's model
class (mm_s2ut/models/mm_s2s_transformer.py:625-700).

``forward`` passes the nine encoder kwargs through ``forward_encoder``, runs the unit decoder on the encoder output
(``self.decoder(prev_output_tokens, encoder_out=encoder_out)``, :693-696) and attaches ``encoder_states`` /
``encoder_padding_mask`` to the decoder's ``extra`` dict when ``return_all_hiddens`` (:697-699).  Parameters carry
fairseq's names (``encoder.*`` / ``decoder.*``) so a reference checkpoint loads with ``load_state_dict``.  Both halves
run on the CUDA kernels (``engine.EncoderEngine`` / ``decoder.UnitDecoderEngine``); generation (incremental decoding,
beam search) is not part of this class.  With fairseq installed the registered model in ``fairseq_glue.py`` is the one
fairseq-train / fairseq-generate instantiate; this class is the stand-alone twin.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from ..decoder import UnitDecoderEngine
from .mm_s2s_transformer import MM_S2STransformerEncoder
from .modules import SelfAttentionParams


class TransformerDecoderLayerParams(nn.Module):
    """fairseq ``TransformerDecoderLayerBase`` parameters (pre-LN, ReLU, encoder attention over kdim = vdim = d)."""

    def __init__(self, embed_dim: int, ffn_dim: int, num_heads: int):
        super().__init__()
        self.self_attn = SelfAttentionParams(embed_dim, num_heads)
        self.self_attn_layer_norm = nn.LayerNorm(embed_dim, eps=1e-5)
        self.encoder_attn = SelfAttentionParams(embed_dim, num_heads)
        self.encoder_attn_layer_norm = nn.LayerNorm(embed_dim, eps=1e-5)
        self.fc1 = nn.Linear(embed_dim, ffn_dim)
        self.fc2 = nn.Linear(ffn_dim, embed_dim)
        self.final_layer_norm = nn.LayerNorm(embed_dim, eps=1e-5)


class TransformerUnitDecoderParams(nn.Module):
    """fairseq ``TransformerUnitDecoder`` parameters: tied input / output embedding, sinusoidal positions (no
    parameters), N layers, final LayerNorm.  ``output_projection.weight`` aliases ``embed_tokens.weight`` as in
    fairseq when ``share_decoder_input_output_embed`` (it appears under both names in a checkpoint)."""

    def __init__(self, vocab: int, embed_dim: int, ffn_dim: int, num_heads: int, num_layers: int, padding_idx: int = 1):
        super().__init__()
        self.padding_idx, self.num_heads = padding_idx, num_heads
        self.embed_tokens = nn.Embedding(vocab, embed_dim, padding_idx=padding_idx)
        nn.init.normal_(self.embed_tokens.weight, mean=0, std=embed_dim ** -0.5)
        nn.init.constant_(self.embed_tokens.weight[padding_idx], 0)
        self.layers = nn.ModuleList(TransformerDecoderLayerParams(embed_dim, ffn_dim, num_heads) for _ in range(num_layers))
        self.layer_norm = nn.LayerNorm(embed_dim, eps=1e-5)
        self.output_projection = nn.Linear(embed_dim, vocab, bias=False)
        self.output_projection.weight = self.embed_tokens.weight


class MM_S2UTTransformerModel(nn.Module):
    def __init__(self, args, target_code_size: int = 1000, build_unused_projections: bool = True):
        super().__init__()
        self.args = args
        self.encoder = MM_S2STransformerEncoder(args, build_unused_projections=build_unused_projections)
        d = int(getattr(args, "decoder_embed_dim", args.encoder_embed_dim))
        self.decoder = TransformerUnitDecoderParams(
            target_code_size + 4, d, int(getattr(args, "decoder_ffn_embed_dim", args.encoder_ffn_embed_dim)),
            int(getattr(args, "decoder_attention_heads", 8)), int(getattr(args, "decoder_layers", 6)))
        self._decoder_engine: Optional[UnitDecoderEngine] = None

    # packed device weights follow the parameters: rebuild after .to() / load_state_dict
    def _apply(self, fn, *a, **k):
        self._decoder_engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._decoder_engine = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def decoder_engine(self) -> UnitDecoderEngine:
        if self._decoder_engine is None:
            dev = self.decoder.embed_tokens.weight.device
            sd = {k: v for k, v in self.decoder.state_dict().items() if not k.startswith("output_projection.")}
            self._decoder_engine = UnitDecoderEngine(sd, self.decoder.num_heads, dev,
                                                     op_dtype=getattr(self.encoder, "op_dtype", torch.bfloat16),
                                                     padding_idx=self.decoder.padding_idx)
        return self._decoder_engine

    # ---- training-step variant (BASELINE configs[2]) -------------------------------------
    def decoder_train_engine(self):
        """``decoder_training.UnitDecoderTrainEngine`` over this model's decoder parameters (its own flat fp32 copy:
        call ``sync_decoder_parameters()`` before saving a checkpoint)."""
        from ..decoder_training import UnitDecoderTrainEngine

        if not isinstance(self._decoder_engine, UnitDecoderTrainEngine):
            dev = self.decoder.embed_tokens.weight.device
            sd = {k: v for k, v in self.decoder.state_dict().items() if not k.startswith("output_projection.")}
            self._decoder_engine = UnitDecoderTrainEngine(sd, self.decoder.num_heads, dev,
                                                          op_dtype=getattr(self.encoder, "op_dtype", torch.bfloat16),
                                                          padding_idx=self.decoder.padding_idx)
            a = self.args      # fairseq TransformerDecoder: --dropout, --attention-dropout, --activation-dropout / --relu-dropout
            self._decoder_engine.dropout_p = float(getattr(a, "dropout", 0.0) or 0.0)
            self._decoder_engine.attention_dropout_p = float(getattr(a, "attention_dropout", 0.0) or 0.0)
            self._decoder_engine.activation_dropout_p = float(getattr(a, "activation_dropout", 0.0) or
                                                              getattr(a, "relu_dropout", 0.0) or 0.0)
        return self._decoder_engine

    @torch.no_grad()
    def sync_decoder_parameters(self) -> None:
        """Write the decoder engine's trained fp32 parameters back into ``self.decoder`` (state_dict / checkpoints)."""
        eng = self._decoder_engine
        if eng is None or not hasattr(eng, "flat_p"):
            return
        params = dict(self.decoder.named_parameters())
        for name in eng.names:
            src = eng.p(name)
            dst = params[name]
            dst.copy_(src[: dst.shape[0]] if name == "embed_tokens.weight" else src.view(dst.shape))

    @torch.no_grad()
    def train_step(self, src_tokens, src_lengths, prev_output_tokens, target, imgs_list=[], img_masks_list=[],
                   label_smoothing: float = 0.2, drop_image: bool = False):
        """Forward + backward of the whole model on the CUDA kernels, no autograd: encoder (activations kept) -> unit
        decoder -> fairseq label-smoothed cross entropy -> decoder backward -> ``d loss / d encoder_out`` -> encoder
        backward.  Returns (loss, nll_loss) as device scalars; gradients are in the two engines' ``flat_g`` (the
        encoder's also as ``param.grad``).  Follow with ``optimizer_step``."""
        eeng, deng = self.encoder.train_engine(), self.decoder_train_engine()
        fuse = bool(imgs_list)
        out = eeng.forward_train(src_tokens, src_lengths, imgs_list if fuse else [], img_masks_list if fuse else [],
                                 drop_image=drop_image)
        mask = out["encoder_padding_mask"][0] if out["encoder_padding_mask"] else None
        deng.forward_train(prev_output_tokens, out["encoder_out"][0], mask)
        loss, nll, d_enc = deng.loss_backward(target, label_smoothing)
        eeng.backward(d_enc)
        return loss, nll

    def optimizer_step(self, lr: float, betas=(0.9, 0.98), eps: float = 1e-8, weight_decay: float = 0.0,
                       clip_norm: float = 0.0, grad_scale: float = 1.0) -> None:
        """fairseq's update on both engines: gradient scaling, clipping by the WHOLE model's norm, Adam."""
        eeng, deng = self.encoder.train_engine(), self.decoder_train_engine()
        deng.grad_norm(grad_scale=grad_scale)
        eeng.adam_step(lr, betas, eps, weight_decay, clip_norm, grad_scale, extra_norm=deng.norm_coef)
        deng.adam_apply(eeng.norm_coef, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)

    def forward_encoder(self, src_tokens, src_lengths, src_audio_path=None, img_path=None, img_tensor=None,
                        imgs_list=[], img_masks_list=[], speaker=None, **kwargs):
        return self.encoder(src_tokens, src_lengths=src_lengths, src_audio_path=src_audio_path, img_path=img_path,
                            img_tensor=img_tensor, imgs_list=imgs_list, img_masks_list=img_masks_list,
                            tgt_speaker=speaker, **kwargs)

    def forward(self, src_tokens, src_lengths, prev_output_tokens, src_audio_path=None, img_path=None, img_tensor=None,
                imgs_list=[], img_masks_list=[], tgt_speaker=None, return_all_hiddens=False,
                **kwargs) -> Tuple[torch.Tensor, Dict[str, List]]:
        if self.training:
            raise NotImplementedError("forward() is the inference path; the training step of the whole model is "
                                      "train_step() / optimizer_step() (DESIGN.md section 9)")
        encoder_out = self.forward_encoder(src_tokens, src_lengths=src_lengths, src_audio_path=src_audio_path,
                                           img_path=img_path, img_tensor=img_tensor, imgs_list=imgs_list,
                                           img_masks_list=img_masks_list, speaker=tgt_speaker,
                                           return_all_hiddens=return_all_hiddens, **kwargs)
        mask = encoder_out["encoder_padding_mask"][0] if encoder_out["encoder_padding_mask"] else None
        logits = self.decoder_engine().forward(prev_output_tokens, encoder_out["encoder_out"][0], mask)
        extra: Dict[str, List] = {"attn": [None], "inner_states": []}
        if return_all_hiddens:
            extra["encoder_states"] = encoder_out["encoder_states"]
            extra["encoder_padding_mask"] = encoder_out["encoder_padding_mask"]
        return logits, extra
