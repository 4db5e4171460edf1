"""Training step of the S2UT unit decoder + label-smoothed cross entropy on the CUDA kernels (SURVEY.md 8f rank 1,
the consumer side of BASELINE configs[2]).

The reference gets this from autograd over fairseq's ``TransformerUnitDecoder`` and its criterion
(mm_s2ut/models/mm_s2s_transformer.py:693-696, mm_s2ut/criterions/speech_to_speech_criterion.py:58-102).  Here

    logits = forward_train(prev_output_tokens, encoder_out [T, B, d], encoder_padding_mask)
    loss, nll, d_encoder_out = loss_backward(target, epsilon)

keeps the activations, forms fairseq's label-smoothed loss, and back-propagates through the tied output projection,
the final LayerNorm, N x (FFN, encoder attention, causal self-attention) and the embedding, filling the gradient of
every decoder parameter and returning ``d loss / d encoder_out`` -- exactly the tensor ``TrainEngine.backward`` takes,
so encoder + decoder + criterion form a complete training step without autograd.

Same building blocks as the encoder's backward (``training.py``): every contraction is ``mm_gemm`` with MN-major /
split-K / (sequence, head) operands; softmax backward from recomputed scores (``causal`` for the self-attention,
encoder key lengths for the encoder attention); ``mm_layernorm_bwd``; plus ``mm_label_smoothed_nll_bwd`` and
``mm_embed_tokens_bwd``.  Parameters / gradients / Adam state / 16-bit operand copies are flat buffers; the embedding is
stored with its rows padded to a multiple of 8 (TMA).  Dropout (``dropout_p``, ``attention_dropout_p``,
``activation_dropout_p``) uses the encoder's counter-based masks (``mm_dropout`` / ``mm_softmax_dropout_bwd``) at
fairseq's sites.  No CPU fallback.
"""
from __future__ import annotations

import os

import math
from typing import Dict, List, Optional, Tuple

import torch

from . import kernels as K
from .decoder import UnitDecoderEngine, _round_up
from .training import TrainEngine, _scope


# dropout sites of the decoder (disjoint from the encoder's: the two engines may share one seed)
DSITE_EMBED = 1 << 20


def dsite_layer(i: int, which: int) -> int:
    """which: 0 after self-attention, 1 self-attention probabilities, 2 after encoder attention, 3 encoder-attention
    probabilities, 4 activation, 5 after fc2."""
    return (1 << 20) + 16 * (i + 1) + which


class UnitDecoderTrainEngine(UnitDecoderEngine):
    _require_cuda = True
    # fairseq TransformerDecoder dropouts (set before forward_train; --dropout, --attention-dropout, --activation-dropout)
    dropout_p = 0.0
    attention_dropout_p = 0.0
    activation_dropout_p = 0.0

    # backward helpers shared with the encoder engine (they only use self.buf / block_n / op_dtype / _ln_blocks)
    _ARENA = TrainEngine._ARENA
    _partials = TrainEngine._partials
    _defer = TrainEngine._defer
    _flush = TrainEngine._flush
    _lnp = TrainEngine._lnp
    _wgrad_mn = TrainEngine._wgrad_mn
    _wgrad_flush = TrainEngine._wgrad_flush
    _wq_key = staticmethod(TrainEngine._wq_key)
    _dgrad_ln_bwd = TrainEngine._dgrad_ln_bwd
    fused_ln_bwd = os.environ.get("MM_FUSED_LN_BWD", "0") != "0"      # see TrainEngine.fused_ln_bwd
    grouped_wgrad = os.environ.get("MM_GROUPED_WGRAD", "1") != "0"     # see TrainEngine.grouped_wgrad
    heads_gemm = os.environ.get("MM_HEADS_GEMM", "1") != "0"
    fused_attn_bwd_onchip = os.environ.get("MM_ATTN_BWD_ONCHIP", "1") != "0"
    _bias_grad = TrainEngine._bias_grad
    _linear_bwd = TrainEngine._linear_bwd
    _ln_param_grads = TrainEngine._ln_param_grads

    def __init__(self, state_dict: Dict[str, torch.Tensor], heads: int, device, op_dtype: torch.dtype = torch.bfloat16,
                 padding_idx: int = 1):
        self.device = torch.device(device)
        if self._require_cuda and self.device.type != "cuda":
            raise RuntimeError("the unit decoder (B200 build) runs only on a CUDA device; there is no CPU fallback")
        self.op_dtype, self.heads, self.padding_idx, self.block_n = op_dtype, heads, padding_idx, 256
        emb = state_dict["embed_tokens.weight"]
        self.vocab, self.d = emb.shape
        if self.d // heads != 64 or self.d % 128 != 0:
            raise NotImplementedError("attention kernels are built for head_dim 64 and d_model % 128 == 0")
        self.vocab_pad = _round_up(self.vocab, 8)
        self.embed_scale = math.sqrt(self.d)
        self.fused_ln = False
        self.fused_attn_bwd = True     # attention backward: scores / dP / softmax backward in one kernel
        self._buf: Dict[Tuple, torch.Tensor] = {}
        self._pos = None
        self.n_layers = 0
        while f"layers.{self.n_layers}.fc1.weight" in state_dict:
            self.n_layers += 1
        self.ffn = state_dict["layers.0.fc1.weight"].shape[0]
        # ---- flat buffers; order keeps q|k|v, k|v and LayerNorm weight|bias adjacent
        names: List[str] = ["embed_tokens.weight"]
        for i in range(self.n_layers):
            p = f"layers.{i}."
            sa, ea = p + "self_attn.", p + "encoder_attn."
            names += [p + "self_attn_layer_norm.weight", p + "self_attn_layer_norm.bias",
                      sa + "q_proj.weight", sa + "k_proj.weight", sa + "v_proj.weight",
                      sa + "q_proj.bias", sa + "k_proj.bias", sa + "v_proj.bias", sa + "out_proj.weight", sa + "out_proj.bias",
                      p + "encoder_attn_layer_norm.weight", p + "encoder_attn_layer_norm.bias",
                      ea + "q_proj.weight", ea + "q_proj.bias", ea + "k_proj.weight", ea + "v_proj.weight",
                      ea + "k_proj.bias", ea + "v_proj.bias", ea + "out_proj.weight", ea + "out_proj.bias",
                      p + "final_layer_norm.weight", p + "final_layer_norm.bias",
                      p + "fc1.weight", p + "fc1.bias", p + "fc2.weight", p + "fc2.bias"]
        names += ["layer_norm.weight", "layer_norm.bias"]
        self._slices: Dict[str, Tuple[int, int, tuple]] = {}
        total = 0
        for n in names:
            shape = tuple(state_dict[n].shape)
            numel = int(torch.Size(shape).numel())
            if n == "embed_tokens.weight":
                shape, numel = (self.vocab_pad, self.d), self.vocab_pad * self.d
            self._slices[n] = (total, numel, shape)
            total += _round_up(numel, 8)
        self.names = names
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=self.device)
        for n in names:
            o, numel, shape = self._slices[n]
            src = state_dict[n].detach().to(self.device, torch.float32).reshape(-1)
            self.flat_p[o:o + src.numel()].copy_(src)
        self.flat_op = torch.empty(total, dtype=op_dtype, device=self.device)
        K.convert(self.flat_p, self.flat_op)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.norm_coef = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._sumsq_partials = torch.zeros(K._lib.load().mm_sumsq_blocks(), dtype=torch.float32, device=self.device)
        self._ln_blocks = K.layernorm_bwd_blocks()
        self.step_count = 0
        self._saved = None
        d = self.d
        self.emb_f32 = self.p("embed_tokens.weight")[: self.vocab]
        self.emb_op = self.op("embed_tokens.weight")
        self.layers = []
        for i in range(self.n_layers):
            p = f"layers.{i}."
            sa, ea = p + "self_attn.", p + "encoder_attn."
            self.layers.append(dict(
                wqkv=self.op(sa + "q_proj.weight", sa + "k_proj.weight", sa + "v_proj.weight").view(3 * d, d),
                bqkv=self.p(sa + "q_proj.bias", sa + "k_proj.bias", sa + "v_proj.bias"),
                wo=self.op(sa + "out_proj.weight"), bo=self.p(sa + "out_proj.bias"),
                ln1=(self.p(p + "self_attn_layer_norm.weight"), self.p(p + "self_attn_layer_norm.bias")),
                wq=self.op(ea + "q_proj.weight"), bq=self.p(ea + "q_proj.bias"),
                wkv=self.op(ea + "k_proj.weight", ea + "v_proj.weight").view(2 * d, d),
                bkv=self.p(ea + "k_proj.bias", ea + "v_proj.bias"),
                wo2=self.op(ea + "out_proj.weight"), bo2=self.p(ea + "out_proj.bias"),
                ln2=(self.p(p + "encoder_attn_layer_norm.weight"), self.p(p + "encoder_attn_layer_norm.bias")),
                w1=self.op(p + "fc1.weight"), b1=self.p(p + "fc1.bias"), w2=self.op(p + "fc2.weight"),
                b2=self.p(p + "fc2.bias"),
                ln3=(self.p(p + "final_layer_norm.weight"), self.p(p + "final_layer_norm.bias")), prefix=p))
        self.ln_out = (self.p("layer_norm.weight"), self.p("layer_norm.bias"))

    # ------------------------------------------------------------------------------------------
    def _span(self, flat: torch.Tensor, names) -> torch.Tensor:
        o0, n0, shape = self._slices[names[0]]
        end = o0 + n0
        for n in names[1:]:
            o, k, _ = self._slices[n]
            assert o == end, "parameters are not adjacent in the flat buffer"
            end = o + k
        t = flat[o0:end]
        return t.view(shape) if len(names) == 1 else t

    def p(self, *names) -> torch.Tensor:
        return self._span(self.flat_p, names)

    def g(self, *names) -> torch.Tensor:
        return self._span(self.flat_g, names)

    def op(self, *names) -> torch.Tensor:
        return self._span(self.flat_op, names)

    def grads(self) -> Dict[str, torch.Tensor]:
        """Gradient of every parameter under its fairseq state-dict name (the embedding without its padding rows)."""
        out = {n: self.g(n) for n in self.names}
        out["embed_tokens.weight"] = out["embed_tokens.weight"][: self.vocab]
        return out

    def buf(self, name: str, shape, dtype, zero: bool = False) -> torch.Tensor:
        key = (name, tuple(shape), dtype)
        t = self._buf.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)
            self._buf[key] = t
        return t

    # ------------------------------------------------------------------------------------------
    # forward (activations kept)
    # ------------------------------------------------------------------------------------------
    def _resid(self, a, a_ld, w, b, k, x_in, x_out, M, site, ln=None, h=None):
        """x_out = x_in + dropout(a w^T + b), then h = LayerNorm_ln(x_out): the fused GEMM + dropout + residual +
        LayerNorm kernel with a separate output (d_model 512), else GEMM (-> mm_dropout(+resid)) -> LayerNorm."""
        d, bn = self.d, self.block_n
        p, _, _, seed, seed_dev = self._drop
        if d == 512 and ln is not None and a_ld == k:
            K.gemm_resid_ln(a, w, b, x_in, ln[0], ln[1], h, x_out=x_out, drop=(p, seed, seed_dev, site) if p > 0 else None)
            return
        if p > 0:
            y = self.buf("t_y", (M, d), torch.float32)
            K.gemm(a0=a, a0_ld=a_ld, rows=M, w=w, n=d, k=k, mode=K.EPI_F32, bias=b, out0=y, out0_ld=d, block_n=bn)
            K.dropout(y, x_out, p, seed, site, resid=x_in, seed_dev=seed_dev)
        else:
            K.gemm(a0=a, a0_ld=a_ld, rows=M, w=w, n=d, k=k, mode=K.EPI_RESID_F32, bias=b, aux0=x_in, aux_ld=d, out0=x_out,
                   out0_ld=d, block_n=bn)
        if ln is not None:
            K.layernorm(x_out, ln[0], ln[1], out_op=h)

    def _attn_fwd(self, q, q_ld, q_bs, k, v, kv_ld, kv_bs, out, B, Lq, Tk, kv_lens, causal, site):
        """Attention forward: the fused kernel, or -- with attention dropout -- scores / softmax+dropout / P V on the
        head-mode GEMMs (the backward pass regenerates the mask from the same site)."""
        d, H, op, bn = self.d, self.heads, self.op_dtype, self.block_n
        _, p_attn, _, seed, seed_dev = self._drop
        BH, Lp, Tp = B * H, _round_up(Lq, 64), _round_up(Tk, 64)
        hd = dict(heads=H, head_stride=64, batches=BH, w_batched=True, block_n=bn)
        S = self.buf("a_S", (BH, Lp, Tp), torch.float32)
        P = self.buf("a_P", (BH, Lp, Tp), op)
        K.gemm(a0=q, a0_ld=q_ld, a0_bs=q_bs, w=k, w_ld=kv_ld, w_bs=kv_bs, out0=S, rows=Lq, n=Tk, k=64, mode=K.EPI_F32,
               out0_ld=Tp, out0_bs=Lp * Tp, a_hm=True, w_hm=True, **hd)
        K.softmax_bwd(S, None, Tp, BH * Lp, Lp, Tk, None, Tp, probs=P, kv_lens=kv_lens, heads=H, valid_rows=Lq,
                      causal=causal, drop_p=p_attn, seed=seed, seed_dev=seed_dev, site=site)
        K.gemm(a0=P, a0_ld=Tp, a0_bs=Lp * Tp, rows=Lq, k=Tk, w=v, w_ld=kv_ld, w_bs=kv_bs, w_mn=True, w_hm=True, n=64,
               mode=K.EPI_OP, out0=out, out0_ld=d, out0_bs=Lq * d, out_hm=True, **hd)

    @torch.no_grad()
    def forward_train(self, prev_output_tokens: torch.Tensor, encoder_out: torch.Tensor,
                      encoder_padding_mask: Optional[torch.Tensor] = None, dropout_seed: Optional[int] = None,
                      dropout_seed_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
        """prev_output_tokens [B, L] int64, encoder_out [T, B, d] fp32 -> logits [B, L, V] fp32 (a view)."""
        dev, d, op, H, bn = self.device, self.d, self.op_dtype, self.heads, self.block_n
        if dropout_seed is None:
            self._drop_calls = getattr(self, "_drop_calls", 0) + 1
            dropout_seed = (torch.initial_seed() + 0x2545F491 * self._drop_calls) & 0x7FFFFFFFFFFFFFFF
        self._drop = (float(self.dropout_p), float(self.attention_dropout_p), float(self.activation_dropout_p),
                      int(dropout_seed), dropout_seed_dev)
        p_drop, p_attn, p_act, seed, seed_dev = self._drop
        tokens = prev_output_tokens.to(dev).contiguous()
        B, L = tokens.shape
        T = encoder_out.shape[0]
        assert encoder_out.shape[1] == B and encoder_out.shape[2] == d
        M = B * L
        enc = encoder_out.to(dev, torch.float32).contiguous()
        enc_lens = None
        if encoder_padding_mask is not None and encoder_padding_mask.numel():
            enc_lens = (~encoder_padding_mask.to(dev)).sum(1).to(torch.int32).contiguous()   # index glue
        # token-major 16-bit copy of the encoder states: [T, B, d] fp32 -> [B, T, d]
        enc_btc = self.buf("enc_btc", (B * T, d), op)
        K.pack_t(enc, rows=T, cols=d, in_ld=B * d, batches=B, in_bs0=d, out_n=enc_btc, n_ld=d, n_bs0=T * d)
        x = self.buf("x0", (M, d), torch.float32)
        K.embed_tokens(tokens, self.padding_idx, self.emb_f32, self.embed_scale, self.pos_table(L + self.padding_idx + 1), x)
        if p_drop > 0:
            K.dropout(x, x, p_drop, seed, DSITE_EMBED, seed_dev=seed_dev)
        scale = 64 ** -0.5
        saved = dict(B=B, L=L, T=T, tokens=tokens, enc_lens=enc_lens, enc_btc=enc_btc, layers=[], drop=self._drop)
        h_out = self.buf("h_out", (M, d), op)
        for i, Lr in enumerate(self.layers):
            s = dict(x0=x, h1=self.buf(f"h1_{i}", (M, d), op), qkv=self.buf(f"qkv_{i}", (M, 3 * d), op),
                     att=self.buf(f"att_{i}", (M, d), op), x1=self.buf(f"x1_{i}", (M, d), torch.float32),
                     h2=self.buf(f"h2_{i}", (M, d), op), q2=self.buf(f"q2_{i}", (M, d), op),
                     kv2=self.buf(f"kv2_{i}", (B * T, 2 * d), op), att2=self.buf(f"att2_{i}", (M, d), op),
                     x2=self.buf(f"x2_{i}", (M, d), torch.float32), h3=self.buf(f"h3_{i}", (M, d), op),
                     f=self.buf(f"f_{i}", (M, self.ffn), op), x3=self.buf(f"x3_{i}", (M, d), torch.float32))
            if i == 0:       # later layers: LN1 was produced by the previous layer's fc2 step
                K.layernorm(s["x0"], Lr["ln1"][0], Lr["ln1"][1], out_op=s["h1"])
            K.gemm(a0=s["h1"], a0_ld=d, rows=M, w=Lr["wqkv"], n=3 * d, k=d, mode=K.EPI_OP, bias=Lr["bqkv"], scale=scale,
                   scale_cols=d, out0=s["qkv"], out0_ld=3 * d, block_n=bn)
            onchip = self.fused_attn_bwd and self.fused_attn_bwd_onchip    # attention dropout inside the attention kernels
            adr = lambda k_: (p_attn, seed, seed_dev, dsite_layer(i, k_)) if p_attn > 0 else None
            if p_attn > 0 and not onchip:
                self._attn_fwd(s["qkv"], 3 * d, L * 3 * d, s["qkv"][:, d:], s["qkv"][:, 2 * d:], 3 * d, L * 3 * d, s["att"],
                               B, L, L, None, True, dsite_layer(i, 1))
            else:
                s["lse1"] = self.buf(f"lse1_{i}", (B, H, L), torch.float32)      # kept for the attention backward
                K.attention(s["qkv"], 0, L, s["qkv"], d, s["qkv"], 2 * d, L, None, B, H, s["att"], causal=True, lse=s["lse1"],
                            drop=adr(1))
            self._resid(s["att"], d, Lr["wo"], Lr["bo"], d, s["x0"], s["x1"], M, dsite_layer(i, 0), Lr["ln2"], s["h2"])
            K.gemm(a0=s["h2"], a0_ld=d, rows=M, w=Lr["wq"], n=d, k=d, mode=K.EPI_OP, bias=Lr["bq"], scale=scale,
                   scale_cols=d, out0=s["q2"], out0_ld=d, block_n=bn)
            K.gemm(a0=enc_btc, a0_ld=d, rows=B * T, w=Lr["wkv"], n=2 * d, k=d, mode=K.EPI_OP, bias=Lr["bkv"],
                   out0=s["kv2"], out0_ld=2 * d, block_n=bn)
            if p_attn > 0 and not onchip:
                self._attn_fwd(s["q2"], d, L * d, s["kv2"], s["kv2"][:, d:], 2 * d, T * 2 * d, s["att2"], B, L, T, enc_lens,
                               False, dsite_layer(i, 3))
            else:
                s["lse2"] = self.buf(f"lse2_{i}", (B, H, L), torch.float32)
                K.attention(s["q2"], 0, L, s["kv2"], 0, s["kv2"], d, T, enc_lens, B, H, s["att2"], lse=s["lse2"], drop=adr(3))
            self._resid(s["att2"], d, Lr["wo2"], Lr["bo2"], d, s["x1"], s["x2"], M, dsite_layer(i, 2), Lr["ln3"], s["h3"])
            K.gemm(a0=s["h3"], a0_ld=d, rows=M, w=Lr["w1"], n=self.ffn, k=d, mode=K.EPI_RELU_OP, bias=Lr["b1"],
                   out0=s["f"], out0_ld=self.ffn, block_n=bn,
                   drop=(p_act, seed, seed_dev, dsite_layer(i, 4)) if p_act > 0 else None)     # activation dropout in the epilogue
            last = i + 1 == self.n_layers
            nxt_ln = self.ln_out if last else self.layers[i + 1]["ln1"]
            nxt_h = h_out if last else self.buf(f"h1_{i + 1}", (M, d), op)
            self._resid(s["f"], self.ffn, Lr["w2"], Lr["b2"], self.ffn, s["x2"], s["x3"], M, dsite_layer(i, 5), nxt_ln, nxt_h)
            saved["layers"].append(s)
            x = s["x3"]
        h = h_out
        logits = self.buf("logits", (M, self.vocab_pad), torch.float32)
        K.gemm(a0=h, a0_ld=d, rows=M, w=self.emb_op, n=self.vocab_pad, k=d, mode=K.EPI_F32, out0=logits,
               out0_ld=self.vocab_pad, block_n=bn)
        saved.update(x_final=x, h_out=h, logits=logits)
        self._saved = saved
        return logits.view(B, L, self.vocab_pad)[:, :, : self.vocab]

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def _attn_bwd(self, q, q_ld, q_bs, k, v, kv_ld, kv_bs, dO, dq, dk, dv, dkv_ld, dkv_bs, B, Lq, Tk, kv_lens, causal,
                  site=0, out=None, lse=None):
        """Attention backward per (sequence, head): q [B][Lq][..] pre-scaled, k / v [B][Tk][..] (column blocks of the
        given tensors), dO [B*Lq, d] -> dq (x head_dim^-0.5), dk, dv written at their heads' column blocks.
        out / lse: the forward attention output and its per-row log-sum-exp -- with them (no attention dropout) the
        scores, dP and the softmax backward are one kernel (mm_attention_bwd_scores)."""
        d, H, op, bn = self.d, self.heads, self.op_dtype, self.block_n
        BH = B * H
        Lp, Tp = _round_up(Lq, 64), _round_up(Tk, 64)
        hd = dict(heads=H, head_stride=64, batches=BH, w_batched=True, block_n=bn)
        _, p_attn, _, seed, seed_dev = self._saved["drop"]
        if lse is not None and self.fused_attn_bwd and self.fused_attn_bwd_onchip:
            # one kernel: S, dP, P, dS never leave the SM (mm_attention_bwd_general: query-tile pairs, fp32 dk / dv partials);
            # the attention-dropout mask of the forward kernel is regenerated from the same site
            scratch = self.buf("a_bwd_scratch", (K.attention_bwd_general_scratch_floats(Tk),), torch.float32)
            K.attention_bwd_general(q, Lq, k, v, Tk, kv_lens, B, H, dO, out, lse, dq, dk, dv, scratch, causal=causal,
                                    drop=(p_attn, seed, seed_dev, site) if p_attn > 0 else None)
            return
        P = self.buf("a_P", (BH, Lp, Tp), op)
        dS = self.buf("a_dS", (BH, Lp, Tp), op)
        if lse is not None and p_attn == 0 and self.fused_attn_bwd:
            K.attention_bwd_scores(q, 0, Lq, k, 0, v, 0, Tk, kv_lens, B, H, dO, out, lse, P, dS, causal=causal)
        else:
            S = self.buf("a_S", (BH, Lp, Tp), torch.float32)
            dP = self.buf("a_dP16", (BH, Lp, Tp), op)      # 16-bit gradient; the scores stay fp32
            sc = dict(rows=Lq, n=Tk, k=64, out0_ld=Tp, out0_bs=Lp * Tp, a_hm=True, w_hm=True, **hd)
            K.gemm(a0=q, a0_ld=q_ld, a0_bs=q_bs, w=k, w_ld=kv_ld, w_bs=kv_bs, out0=S, mode=K.EPI_F32, **sc)
            K.gemm(a0=dO, a0_ld=d, a0_bs=Lq * d, w=v, w_ld=kv_ld, w_bs=kv_bs, out0=dP, mode=K.EPI_OP, **sc)
            K.softmax_bwd(S, dP, Tp, BH * Lp, Lp, Tk, dS, Tp, probs=P, kv_lens=kv_lens, heads=H, valid_rows=Lq,
                          causal=causal, drop_p=p_attn, seed=seed, seed_dev=seed_dev, site=site)
        if self.heads_gemm:     # 128 x 64 tiles per (sequence, head): no wasted columns (see TrainEngine.heads_gemm)
            hg = dict(a_ld=Tp, a_bs=Lp * Tp, batch=B, heads=H)
            K.heads_gemm(P, transposed=True, w=dO, w_ld=d, w_bs=Lq * d, out=dv, out_ld=dkv_ld, out_bs=dkv_bs, rows=Tk,
                         k=Lq, **hg)
            K.heads_gemm(dS, transposed=True, w=q, w_ld=q_ld, w_bs=q_bs, out=dk, out_ld=dkv_ld, out_bs=dkv_bs, rows=Tk,
                         k=Lq, **hg)
            K.heads_gemm(dS, transposed=False, w=k, w_ld=kv_ld, w_bs=kv_bs, out=dq, out_ld=q_ld, out_bs=q_bs, rows=Lq,
                         k=Tk, scale=64 ** -0.5, **hg)
            return
        og = dict(n=64, mode=K.EPI_OP, out_hm=True, w_mn=True, w_hm=True, a0_ld=Tp, a0_bs=Lp * Tp, **hd)
        K.gemm(a0=P, a_mn=True, rows=Tk, k=Lq, w=dO, w_ld=d, w_bs=Lq * d, out0=dv, out0_ld=dkv_ld, out0_bs=dkv_bs, **og)
        K.gemm(a0=dS, a_mn=True, rows=Tk, k=Lq, w=q, w_ld=q_ld, w_bs=q_bs, out0=dk, out0_ld=dkv_ld, out0_bs=dkv_bs, **og)
        K.gemm(a0=dS, rows=Lq, k=Tk, w=k, w_ld=kv_ld, w_bs=kv_bs, out0=dq, out0_ld=q_ld, out0_bs=q_bs, scale=64 ** -0.5,
               scale_cols=64, **og)

    @torch.no_grad()
    def loss_backward(self, target: torch.Tensor, epsilon: float, grad_scale: float = 1.0, accumulate: bool = False):
        """fairseq label-smoothed cross entropy of the last forward_train() against target [B, L] (padding_idx ignored),
        then the backward pass.  Returns (loss, nll_loss, d loss / d encoder_out [T, B, d] fp32), the gradient scaled by
        grad_scale; parameter gradients land in ``flat_g`` (see ``grads()``)."""
        sv = self._saved
        if sv is None:
            raise RuntimeError("loss_backward() needs a preceding forward_train()")
        B, L, T, d, op, bn, ffn = sv["B"], sv["L"], sv["T"], self.d, self.op_dtype, self.block_n, self.ffn
        M, Vp = B * L, self.vocab_pad
        target = target.to(self.device).contiguous().view(-1)
        logits = sv["logits"]
        loss, nll = K.label_smoothed_nll(logits, self.vocab, target, self.padding_idx, epsilon)
        dlog = self.buf("dlogits", (M, Vp), op)
        K.label_smoothed_nll_bwd(logits, self.vocab, target, self.padding_idx, epsilon, dlog, grad_scale)
        dh = self.buf("b_dh", (M, d), torch.float32)
        g = self.buf("b_g", (M, d), torch.float32)
        g_op = self.buf("b_g_op", (M, d), op)
        p_drop, p_attn, p_act, seed, seed_dev = sv["drop"]

        def drb(site):      # the gradient entering a dropped branch is g o mask / (1 - p) (the residual path keeps g itself):
            return (p_drop, seed, seed_dev, site) if p_drop > 0 else None      # the LayerNorm backward before it emits that 16-bit copy
        # ---- tied output projection: logits = h E^T
        with _scope("out_proj"):
            self._wgrad_mn(dlog, Vp, [(sv["h_out"], d, d)], M, Vp, self.g("embed_tokens.weight"), accumulate)
            K.gemm(a0=dlog, a0_ld=Vp, rows=M, w=self.emb_op, w_ld=d, w_mn=True, n=d, k=Vp, mode=K.EPI_F32, out0=dh,
                   out0_ld=d, block_n=bn)
            lnp = self._lnp()
            K.layernorm_bwd(sv["x_final"], self.ln_out[0], dh, lnp, dx=g, dx_op=g_op, drop=drb(dsite_layer(self.n_layers - 1, 5)))
            self._ln_param_grads(lnp, d, self.g("layer_norm.weight", "layer_norm.bias"), accumulate)
            self._flush()
        denc = self.buf("denc_btc", (B * T, d), torch.float32)
        first_kv = True

        grouped = self.grouped_wgrad     # queued weight gradients: every layer keeps its own gradient buffers

        for i in reversed(range(self.n_layers)):
            s, Lr = sv["layers"][i], self.layers[i]
            p = Lr["prefix"]
            sa, ea = p + "self_attn.", p + "encoder_attn."
            tag = f"@{i}" if grouped else ""
            with _scope("dec_layer"):
                # ---- FFN
                gm = g_op           # already masked for this branch's site by the LayerNorm backward that produced it
                self._linear_bwd(gm, d, s["f"], M, d, ffn, self.g(p + "fc2.weight"), self.g(p + "fc2.bias"), accumulate)
                dF = self.buf("b_dF" + tag, (M, ffn), op)
                # ReLU (and activation-dropout) mask in the dgrad's epilogue: the kept activation is > 0 exactly where ReLU passed
                # and the dropout kept it; the surviving gradient is scaled by 1 / (1 - p_act)
                K.gemm(a0=gm, a0_ld=d, rows=M, w=Lr["w2"], w_ld=ffn, w_mn=True, n=ffn, k=d, mode=K.EPI_MASK_OP, out0=dF,
                       out0_ld=ffn, aux0=s["f"], aux_ld=ffn, scale=1.0 / (1.0 - p_act), block_n=bn)
                self._linear_bwd(dF, ffn, s["h3"], M, ffn, d, self.g(p + "fc1.weight"), self.g(p + "fc1.bias"), accumulate)
                if grouped:
                    g_op = self.buf("b_g_op2" + tag, (M, d), op)
                self._dgrad_ln_bwd(dF, Lr["w1"], s["x2"], Lr["ln3"][0], g, g_op,
                                   self.g(p + "final_layer_norm.weight", p + "final_layer_norm.bias"), accumulate,
                                   drb(dsite_layer(i, 2)))
                # ---- encoder attention
                gm = g_op
                self._linear_bwd(gm, d, s["att2"], M, d, d, self.g(ea + "out_proj.weight"), self.g(ea + "out_proj.bias"),
                                 accumulate)
                datt = self.buf("b_datt", (M, d), op)
                K.gemm(a0=gm, a0_ld=d, rows=M, w=Lr["wo2"], w_ld=d, w_mn=True, n=d, k=d, mode=K.EPI_OP, out0=datt,
                       out0_ld=d, block_n=bn)
                dq2 = self.buf("b_dq2" + tag, (M, d), op)
                dkv2 = self.buf("b_dkv2" + tag, (B * T, 2 * d), op)
                kv2 = s["kv2"]
                self._attn_bwd(s["q2"], d, L * d, kv2, kv2[:, d:], 2 * d, T * 2 * d, datt, dq2, dkv2, dkv2[:, d:], 2 * d,
                               T * 2 * d, B, L, T, sv["enc_lens"], False, dsite_layer(i, 3), out=s["att2"],
                               lse=s.get("lse2"))
                self._linear_bwd(dq2, d, s["h2"], M, d, d, self.g(ea + "q_proj.weight"), self.g(ea + "q_proj.bias"),
                                 accumulate)
                self._linear_bwd(dkv2, 2 * d, sv["enc_btc"], B * T, 2 * d, d, self.g(ea + "k_proj.weight", ea + "v_proj.weight"),
                                 self.g(ea + "k_proj.bias", ea + "v_proj.bias"), accumulate)
                if first_kv:   # gradient of the encoder states, summed over the layers
                    K.gemm(a0=dkv2, a0_ld=2 * d, rows=B * T, w=Lr["wkv"], w_ld=d, w_mn=True, n=d, k=2 * d, mode=K.EPI_F32,
                           out0=denc, out0_ld=d, block_n=bn)
                    first_kv = False
                else:
                    K.gemm(a0=dkv2, a0_ld=2 * d, rows=B * T, w=Lr["wkv"], w_ld=d, w_mn=True, n=d, k=2 * d,
                           mode=K.EPI_RESID_F32, aux0=denc, aux_ld=d, out0=denc, out0_ld=d, block_n=bn)
                if grouped:
                    g_op = self.buf("b_g_op1" + tag, (M, d), op)
                self._dgrad_ln_bwd(dq2, Lr["wq"], s["x1"], Lr["ln2"][0], g, g_op,
                                   self.g(p + "encoder_attn_layer_norm.weight", p + "encoder_attn_layer_norm.bias"),
                                   accumulate, drb(dsite_layer(i, 0)))
                # ---- causal self-attention
                gm = g_op
                self._linear_bwd(gm, d, s["att"], M, d, d, self.g(sa + "out_proj.weight"), self.g(sa + "out_proj.bias"),
                                 accumulate)
                K.gemm(a0=gm, a0_ld=d, rows=M, w=Lr["wo"], w_ld=d, w_mn=True, n=d, k=d, mode=K.EPI_OP, out0=datt,
                       out0_ld=d, block_n=bn)
                dqkv = self.buf("b_dqkv" + tag, (M, 3 * d), op)
                qkv = s["qkv"]
                self._attn_bwd(qkv, 3 * d, L * 3 * d, qkv[:, d:], qkv[:, 2 * d:], 3 * d, L * 3 * d, datt, dqkv, dqkv[:, d:],
                               dqkv[:, 2 * d:], 3 * d, L * 3 * d, B, L, L, None, True, dsite_layer(i, 1), out=s["att"],
                               lse=s.get("lse1"))
                self._linear_bwd(dqkv, 3 * d, s["h1"], M, 3 * d, d,
                                 self.g(sa + "q_proj.weight", sa + "k_proj.weight", sa + "v_proj.weight"),
                                 self.g(sa + "q_proj.bias", sa + "k_proj.bias", sa + "v_proj.bias"), accumulate)
                if grouped:
                    g_op = self.buf("b_g_op0" + tag, (M, d), op)
                self._dgrad_ln_bwd(dqkv, Lr["wqkv"], s["x0"], Lr["ln1"][0], g, g_op,
                                   self.g(p + "self_attn_layer_norm.weight", p + "self_attn_layer_norm.bias"), accumulate,
                                   drb(dsite_layer(i - 1, 5)) if i > 0 else None)
                if not grouped:
                    self._flush()   # the layer's deferred reductions in one launch (per 16)
        with _scope("dec_layer"):
            self._wgrad_flush()     # every queued weight gradient (the embedding's wgrad must precede its scatter-add below)
            self._flush()           # pooled weight gradients: only the LayerNorm-parameter partials are left, all layers at once
        # ---- embedding: x0 = dropout(sqrt(d) * E[tokens] + positions)
        if p_drop > 0:
            K.dropout(g, g, p_drop, seed, DSITE_EMBED, seed_dev=seed_dev)
        K.embed_tokens_bwd(sv["tokens"].view(-1), self.padding_idx, g, self.embed_scale, self.g("embed_tokens.weight"))
        d_enc = torch.empty(T, B, d, dtype=torch.float32, device=self.device)
        K.tbc_to_btc(denc, T, B, d, d_enc)       # [B, T, d] -> [T, B, d]: the same index swap with the roles exchanged
        return loss, nll, d_enc

    def grad_norm(self, grad_scale: float = 1.0, dev_hyper: bool = False) -> torch.Tensor:
        """norm_coef[0] = ||grad_scale * flat_g|| on the device (no clipping): the decoder's share of the model's joint
        gradient norm, handed to the encoder engine's ``adam_step(extra_norm=...)``."""
        K.grad_clip_coef(self.flat_g, grad_scale, 0.0, self._sumsq_partials, self.norm_coef, dev_hyper=dev_hyper)
        return self.norm_coef

    def adam_apply(self, norm_coef: torch.Tensor, lr: float = 0.0, betas=(0.9, 0.98), eps: float = 1e-8,
                   weight_decay: float = 0.0, step: Optional[int] = None) -> None:
        """Adam with the clip multiplier (and, for step == 0, the per-step hyper-parameters) of ``norm_coef`` -- the
        ENCODER engine's array when the two engines are clipped jointly (fairseq clips the whole model's norm)."""
        if step is None:
            self.step_count += 1
            step = self.step_count
        K.adam(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, lr=lr, betas=betas, eps=eps,
               weight_decay=weight_decay, step=step, norm_coef=norm_coef, param_op=self.flat_op)

    def adam_step(self, lr: float, betas=(0.9, 0.98), eps: float = 1e-8, weight_decay: float = 0.0,
                  clip_norm: float = 0.0, grad_scale: float = 1.0) -> None:
        self.step_count += 1
        K.grad_clip_coef(self.flat_g, grad_scale, clip_norm, self._sumsq_partials, self.norm_coef)
        K.adam(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, lr=lr, betas=betas, eps=eps,
               weight_decay=weight_decay, step=self.step_count, norm_coef=self.norm_coef, param_op=self.flat_op)
