"""Teacher-forced forward of the S2UT unit decoder on the CUDA kernels (forward only).

fairseq ``TransformerUnitDecoder`` as configured by ``s2ut_architecture_base`` (reference call site
mm_s2ut/models/mm_s2s_transformer.py:693-696, ``self.decoder(prev_output_tokens, encoder_out=encoder_out)``):
token embedding (V = target_code_size + 4, padding_idx 1, scale sqrt(d)) + sinusoidal positions, N pre-LN layers
(causal self-attention, encoder attention with the encoder's key-padding mask, ReLU FFN), final LayerNorm, output
projection tied to the embedding, n_frames_per_step = 1.  It is the first consumer of the fused encoder states.

Kernels: ``mm_embed_tokens``; ``mm_gemm`` (QKV, q / K|V of the encoder attention, fc1, tied output projection);
``mm_attention`` (causal self-attention, encoder attention); ``mm_gemm_resid_ln`` (out_proj / fc2 + residual + the
next LayerNorm, d_model = 512) or ``mm_gemm`` RESID_F32 + ``mm_layernorm``.  Parameters come in fairseq's state-dict
naming (``layers.{i}.self_attn.q_proj.weight`` ...).  Not built: incremental (beam-search) decoding, the criterion,
multi-frame (n_frames_per_step > 1) heads, dropout in training mode.  There is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from . import kernels as K
from .models.modules import SinusoidalPositionalEmbedding


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class UnitDecoderEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], heads: int, device, op_dtype: torch.dtype = torch.bfloat16,
                 padding_idx: int = 1):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the unit decoder (B200 build) runs only on a CUDA device; there is no CPU fallback")
        self.op_dtype, self.heads, self.padding_idx = op_dtype, heads, padding_idx
        emb = state_dict["embed_tokens.weight"]
        self.vocab, self.d = emb.shape
        if self.d // heads != 64 or self.d % 128 != 0:
            raise NotImplementedError("attention kernels are built for head_dim 64 and d_model % 128 == 0")
        self.fused_ln = self.d == 512
        self.embed_scale = math.sqrt(self.d)
        self._buf: Dict[Tuple, torch.Tensor] = {}
        self._pos: Optional[torch.Tensor] = None
        sd = state_dict
        f32, op = self._f32, self._op
        self.emb_f32 = f32(emb)
        self.vocab_pad = _round_up(self.vocab, 8)          # TMA rows of the fp32 logits must be 16-byte multiples
        emb_pad = torch.zeros(self.vocab_pad, self.d)
        emb_pad[: self.vocab] = emb.detach().float().cpu()
        self.emb_op = op(emb_pad)
        self.layers = []
        i = 0
        while f"layers.{i}.fc1.weight" in sd:
            p = f"layers.{i}."
            sa, ea = p + "self_attn.", p + "encoder_attn."
            self.layers.append(dict(
                wqkv=op(torch.cat([sd[sa + "q_proj.weight"], sd[sa + "k_proj.weight"], sd[sa + "v_proj.weight"]], 0)),
                bqkv=f32(torch.cat([sd[sa + "q_proj.bias"], sd[sa + "k_proj.bias"], sd[sa + "v_proj.bias"]], 0)),
                wo=op(sd[sa + "out_proj.weight"]), bo=f32(sd[sa + "out_proj.bias"]),
                ln1=(f32(sd[p + "self_attn_layer_norm.weight"]), f32(sd[p + "self_attn_layer_norm.bias"])),
                wq=op(sd[ea + "q_proj.weight"]), bq=f32(sd[ea + "q_proj.bias"]),
                wkv=op(torch.cat([sd[ea + "k_proj.weight"], sd[ea + "v_proj.weight"]], 0)),
                bkv=f32(torch.cat([sd[ea + "k_proj.bias"], sd[ea + "v_proj.bias"]], 0)),
                wo2=op(sd[ea + "out_proj.weight"]), bo2=f32(sd[ea + "out_proj.bias"]),
                ln2=(f32(sd[p + "encoder_attn_layer_norm.weight"]), f32(sd[p + "encoder_attn_layer_norm.bias"])),
                w1=op(sd[p + "fc1.weight"]), b1=f32(sd[p + "fc1.bias"]),
                w2=op(sd[p + "fc2.weight"]), b2=f32(sd[p + "fc2.bias"]),
                ln3=(f32(sd[p + "final_layer_norm.weight"]), f32(sd[p + "final_layer_norm.bias"]))))
            i += 1
        self.ffn = self.layers[0]["w1"].shape[0]
        self.ln_out = (f32(sd["layer_norm.weight"]), f32(sd["layer_norm.bias"]))

    # ------------------------------------------------------------------------------------------
    def _op(self, w: torch.Tensor) -> torch.Tensor:
        w = w.detach().to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty(w.shape, dtype=self.op_dtype, device=self.device)
        K.convert(w, out)
        return out

    def _f32(self, w: torch.Tensor) -> torch.Tensor:
        return w.detach().to(device=self.device, dtype=torch.float32).contiguous()

    def buf(self, name: str, shape, dtype) -> torch.Tensor:
        key = (name, tuple(shape), dtype)
        t = self._buf.get(key)
        if t is None:
            t = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._buf[key] = t
        return t

    def pos_table(self, n_rows: int) -> torch.Tensor:
        if self._pos is None or self._pos.shape[0] < n_rows:
            n = max(n_rows, 1024)
            self._pos = SinusoidalPositionalEmbedding.get_embedding(n, self.d, self.padding_idx).to(self.device).contiguous()
        return self._pos

    def _proj_resid_ln(self, a, w, b, x, ln, h):
        """x += a w^T + b ;  h = LayerNorm_ln(x)  (one kernel when d_model == 512)."""
        d, M = self.d, x.shape[0]
        if self.fused_ln:
            K.gemm_resid_ln(a, w, b, x, ln[0], ln[1], h)
        else:
            K.gemm(a0=a, a0_ld=a.shape[1], rows=M, w=w, n=d, k=a.shape[1], mode=K.EPI_RESID_F32, bias=b, aux0=x, aux_ld=d,
                   out0=x, out0_ld=d)
            K.layernorm(x, ln[0], ln[1], out_op=h)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, prev_output_tokens: torch.Tensor, encoder_out: torch.Tensor,
                encoder_padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """prev_output_tokens [B, L] int64, encoder_out [T, B, d] fp32 (the encoder's output layout),
        encoder_padding_mask [B, T] bool (True = padded; trailing padding) -> logits [B, L, V] fp32."""
        dev, d, op, H = self.device, self.d, self.op_dtype, self.heads
        tokens = prev_output_tokens.to(dev).contiguous()
        B, L = tokens.shape
        T = encoder_out.shape[0]
        assert encoder_out.shape[1] == B and encoder_out.shape[2] == d
        M = B * L
        enc = encoder_out.to(dev, torch.float32).contiguous()
        if encoder_padding_mask is not None and encoder_padding_mask.numel():
            enc_lens = (~encoder_padding_mask.to(dev)).sum(1).to(torch.int32).contiguous()
        else:
            enc_lens = None
        x = self.buf("x", (M, d), torch.float32)
        h = self.buf("h", (M, d), op)
        K.embed_tokens(tokens, self.padding_idx, self.emb_f32, self.embed_scale, self.pos_table(L + self.padding_idx + 1), x)
        K.layernorm(x, self.layers[0]["ln1"][0], self.layers[0]["ln1"][1], out_op=h)
        enc_op = self.buf("enc_op", (T * B, d), op)              # [T, B, d] 16-bit copy of the encoder states
        K.convert(enc.view(T * B, d), enc_op)
        qkv = self.buf("qkv", (M, 3 * d), op)
        att = self.buf("att", (M, d), op)
        q2 = self.buf("q2", (M, d), op)
        kv2 = self.buf("kv2", (B * T, 2 * d), op)
        f = self.buf("ffn", (M, self.ffn), op)
        scale = 64 ** -0.5
        for i, Lr in enumerate(self.layers):
            nxt = self.layers[i + 1]["ln1"] if i + 1 < len(self.layers) else self.ln_out
            # causal self-attention
            K.gemm(a0=h, a0_ld=d, rows=M, w=Lr["wqkv"], n=3 * d, k=d, mode=K.EPI_OP, bias=Lr["bqkv"], scale=scale,
                   scale_cols=d, out0=qkv, out0_ld=3 * d)
            K.attention(qkv, 0, L, qkv, d, qkv, 2 * d, L, None, B, H, att, causal=True)
            self._proj_resid_ln(att, Lr["wo"], Lr["bo"], x, Lr["ln2"], h)
            # encoder attention: q from the decoder, K|V from the [T, B, d] encoder states (batched strided A operand)
            K.gemm(a0=h, a0_ld=d, rows=M, w=Lr["wq"], n=d, k=d, mode=K.EPI_OP, bias=Lr["bq"], scale=scale, scale_cols=d,
                   out0=q2, out0_ld=d)
            K.gemm(a0=enc_op, a0_ld=B * d, a0_bs=d, rows=T, batches=B, w=Lr["wkv"], n=2 * d, k=d, mode=K.EPI_OP,
                   bias=Lr["bkv"], out0=kv2, out0_ld=2 * d, out0_bs=T * 2 * d)
            K.attention(q2, 0, L, kv2, 0, kv2, d, T, enc_lens, B, H, att)
            self._proj_resid_ln(att, Lr["wo2"], Lr["bo2"], x, Lr["ln3"], h)
            # feed-forward
            K.gemm(a0=h, a0_ld=d, rows=M, w=Lr["w1"], n=self.ffn, k=d, mode=K.EPI_RELU_OP, bias=Lr["b1"], out0=f,
                   out0_ld=self.ffn)
            self._proj_resid_ln(f, Lr["w2"], Lr["b2"], x, nxt, h)
        logits = torch.empty(M, self.vocab_pad, dtype=torch.float32, device=dev)
        K.gemm(a0=h, a0_ld=d, rows=M, w=self.emb_op, n=self.vocab_pad, k=d, mode=K.EPI_F32, out0=logits,
               out0_ld=self.vocab_pad)
        return logits.view(B, L, self.vocab_pad)[:, :, : self.vocab]
