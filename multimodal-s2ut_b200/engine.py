"""Device-side execution of the fbank -> fused-encoder hot path.

``EncoderEngine`` turns the fp32 master parameters of ``MM_S2STransformerEncoder`` into packed 16-bit
operand tensors once (per device / dtype), owns the cached activation workspaces in HBM, and issues the
kernel sequence through the C ABI on the current CUDA stream:

  waveform -> fbank (+CMVN partial sums) -> CMVN apply / pad -> Conv1d+GLU -> Conv1d+GLU * sqrt(d) + pos
  -> L x [LN -> QKV GEMM (q scaled, V^T) -> self-attention -> out_proj + residual -> LN -> fc1+ReLU ->
          fc2 + residual] -> final LN
  -> per image type: LN(image) -> K|V projection (V^T) -> q projection -> scores GEMM -> softmax ->
     P V GEMM -> proj -> gate GEMM (sigmoid gate + mix, stored T x B x C)

HBM layout: activations are token-major ``[B, T, C]`` (utterance-contiguous, so attention tiles are plain
TMA boxes); the residual stream is fp32, every GEMM operand 16-bit (bf16 by default); conv inputs carry two
zero frames in front of each utterance so a stride-2 k=5 window is one contiguous row of a strided view.
There is no PyTorch arithmetic on this path (torch supplies memory, streams and trivial index/mask glue).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from . import kernels as K
from .feature_store import StoredImages
from .models.modules import SinusoidalPositionalEmbedding


def _even(n: int) -> int:
    return n + (n & 1)


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def num_frames(n_samples: int) -> int:
    return 0 if n_samples < 400 else 1 + (n_samples - 400) // 160


def sub_len(n: int) -> int:
    return (n - 1) // 2 + 1 if n > 0 else 0


class EncoderEngine:
    _require_cuda = True   # the CPU test-suite checks the host logic against emulated kernels (tests/_emul.py)

    def __init__(self, enc, op_dtype: Optional[torch.dtype] = None, block_n: int = 256):
        p = next(enc.parameters())
        if self._require_cuda and not p.is_cuda:
            raise RuntimeError(
                "mm_s2ut_transformer (B200 build) runs only on a CUDA device: move the model with .cuda(); "
                "there is no CPU fallback")
        self.device = p.device
        self.op_dtype = op_dtype or getattr(enc, "op_dtype", torch.bfloat16)
        self.block_n = block_n
        self.d = enc.embed_dim
        self.heads = enc.num_heads
        self.ffn = enc.ffn_dim
        self.n_layers = enc.num_layers
        if self.d // self.heads != 64 or self.d % 128 != 0:
            raise NotImplementedError("self-attention kernel is built for head_dim 64 and d_model % 128 == 0")
        self.embed_scale = float(enc.embed_scale)
        self.fused_ln = self.d == 512 and getattr(enc, "fuse_layernorm", True)   # mm_gemm_resid_ln needs n == 512
        # flash-style speech -> image attention (mm_cross_attention) needs 256-column blocks of the output
        self.fused_xattn = self.d % 256 == 0 and getattr(enc, "fuse_cross_attention", True)
        self.enc = enc
        self._buf: Dict[Tuple, torch.Tensor] = {}
        self._pos: Optional[torch.Tensor] = None
        self.fbank_tables = K.fbank_tables(self.device)
        self._pack()

    # ------------------------------------------------------------------------------------------
    # weights
    # ------------------------------------------------------------------------------------------
    def _op(self, w: torch.Tensor) -> torch.Tensor:
        w = w.detach().to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty(w.shape, dtype=self.op_dtype, device=self.device)
        K.convert(w, out)
        return out

    @staticmethod
    def _f32(w: torch.Tensor) -> torch.Tensor:
        return w.detach().float().contiguous()

    def _glu_perm(self, n: int) -> torch.Tensor:
        """Row order that puts each GLU value column and its gate column in the same BN-wide tile."""
        cache = self.__dict__.setdefault("_perm_cache", {})
        if n not in cache:
            cache[n] = self._glu_perm_build(n)
        return cache[n]

    def _glu_perm_build(self, n: int) -> torch.Tensor:
        bn, half = self.block_n, n // 2
        idx = []
        for t in range(n // bn):
            a = torch.arange(t * bn // 2, (t + 1) * bn // 2)
            idx += [a, half + a]
        return torch.cat(idx).to(self.device)

    def _pack_conv(self) -> None:
        enc = self.enc
        convs = enc.subsample.conv_layers
        if len(convs) != 2 or any(c.kernel_size[0] != 5 for c in convs):
            raise NotImplementedError("Conv1dSubsampler kernels are built for conv_kernel_sizes='5,5'")
        self.conv = []
        for i, c in enumerate(convs):
            cout, cin, k = c.weight.shape
            if cout % self.block_n or (cin * 2) % 16:
                raise NotImplementedError("conv channel counts must be multiples of the GEMM tile")
            perm = self._glu_perm(cout)
            w = c.weight.detach().float().permute(0, 2, 1).reshape(cout, k * cin)[perm]   # [n, tap*cin + ci]
            # persistent buffers, rewritten in place: a captured CUDA graph keeps reading the same addresses after
            # the training engine refreshes them
            w_op = self.buf(f"conv_w_op{i}", (cout, k * cin), self.op_dtype)
            K.convert(w.contiguous(), w_op)
            b = self.buf(f"conv_b{i}", (cout,), torch.float32)
            b.copy_(c.bias.detach().float()[perm])
            self.conv.append(dict(w=w_op, b=b, cin=cin, cout=cout, k=k))

    def _pack(self) -> None:
        enc = self.enc
        self._pack_conv()
        self.layers = []
        for L in enc.transformer_layers:
            a = L.self_attn
            wqkv = torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0)
            bqkv = torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0)
            self.layers.append(dict(
                ln1_g=self._f32(L.self_attn_layer_norm.weight), ln1_b=self._f32(L.self_attn_layer_norm.bias),
                wqkv=self._op(wqkv), bqkv=self._f32(bqkv),
                wo=self._op(a.out_proj.weight), bo=self._f32(a.out_proj.bias),
                ln2_g=self._f32(L.final_layer_norm.weight), ln2_b=self._f32(L.final_layer_norm.bias),
                w1=self._op(L.fc1.weight), b1=self._f32(L.fc1.bias),
                w2=self._op(L.fc2.weight), b2=self._f32(L.fc2.bias)))
        self.ln_g, self.ln_b = self._f32(enc.layer_norm.weight), self._f32(enc.layer_norm.bias)
        self.fusion = []
        if enc.multimodal_translation_flag and enc.multimodal_attention_type is not None:
            dims = list(enc.mm_config.image_feat_dim)
            for j, dk in enumerate(dims):
                if enc.multimodal_attention_type == "selective_attention":
                    s = enc.selective_attns[j]
                    f = dict(wq=self._op(s.q_proj.weight), bq=self._f32(s.q_proj.bias),
                             wkv=self._op(torch.cat([s.k_proj.weight, s.v_proj.weight], 0)),
                             bkv=self._f32(torch.cat([s.k_proj.bias, s.v_proj.bias], 0)),
                             wp=self._op(s.proj.weight), bp=self._f32(s.proj.bias), bias_kv=None)
                else:  # multimodal_attention: nn.MultiheadAttention packing, one learned extra key/value
                    m = enc.multimodal_attns[j]
                    bq, bk, bv = m.in_proj_bias.detach().chunk(3)
                    f = dict(wq=self._op(m.q_proj_weight), bq=self._f32(bq),
                             wkv=self._op(torch.cat([m.k_proj_weight, m.v_proj_weight], 0)),
                             bkv=self._f32(torch.cat([bk, bv], 0)),
                             wp=self._op(m.out_proj.weight), bp=self._f32(m.out_proj.bias),
                             bias_kv=(self._op(m.bias_k.reshape(-1)), self._op(m.bias_v.reshape(-1))))
                g = enc.gate_denses[j]
                f.update(wg=self._op(g.weight), bg=self._f32(g.bias), dk=dk)
                self.fusion.append(f)
            pn = enc.image_pre_norm_module
            self.img_ln = None if isinstance(pn, torch.nn.Identity) else (self._f32(pn.weight), self._f32(pn.bias))

    # ------------------------------------------------------------------------------------------
    # workspaces
    # ------------------------------------------------------------------------------------------
    def buf(self, name: str, shape, dtype, zero: bool = False) -> torch.Tensor:
        key = (name, tuple(shape), dtype)
        t = self._buf.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)
            self._buf[key] = t
        return t

    def pos_table(self, n_rows: int) -> torch.Tensor:
        if self._pos is None or self._pos.shape[0] < n_rows:
            n = max(n_rows, 1024)
            self._pos = SinusoidalPositionalEmbedding.get_embedding(n, self.d, 1).to(self.device).contiguous()
        return self._pos

    # ------------------------------------------------------------------------------------------
    # stages
    # ------------------------------------------------------------------------------------------
    def frontend(self, src_tokens: torch.Tensor, src_lengths: torch.Tensor, want_features: bool = False,
                 specaug=None):
        """Raw waveform [B, N] (or features [B, m, 80]) -> conv-ready operand buffer x1 [B, m_alloc, 80].
        specaug = (mask table int32 [B, 2 (nf + nt)], nf, nt, mask_value): SpecAugment fused into the CMVN pass."""
        dev, op = self.device, self.op_dtype
        src_tokens = src_tokens.to(dev, non_blocking=True)
        lens = src_lengths.to(device=dev, dtype=torch.int64, non_blocking=True)
        B = src_tokens.shape[0]
        feats_norm = None
        if src_tokens.dim() == 2:
            if src_tokens.dtype not in (torch.float32, torch.int16):
                raise TypeError("raw waveform input must be float32 in int16 range (already x 2**15) or int16 PCM")
            wav = src_tokens if src_tokens.stride(1) == 1 else src_tokens.contiguous()
            m = num_frames(wav.shape[1])
            if m <= 0:
                raise ValueError("waveform shorter than one 25 ms frame")
            raw = self.buf("fbank_raw", (B, m, 80), torch.float32)
            stats = self.buf("cmvn_mean_std", (B, 2, 80), torch.float32)
            K.fbank(wav, lens, raw, self.fbank_tables)
            K.cmvn_stats(raw, lens, True, stats)
            is_samples = True
        elif src_tokens.dim() == 3 and src_tokens.shape[2] == 80:
            raw, stats, m, is_samples = src_tokens.float().contiguous(), None, src_tokens.shape[1], False
        else:
            raise ValueError("src_tokens must be a [B, N] waveform or [B, T, 80] features")
        m_alloc = _even(m + 4)
        x1 = self.buf("x1", (B, m_alloc, 80), op)
        if want_features:
            feats_norm = torch.empty(B, m, 80, dtype=torch.float32, device=dev)
        if specaug is not None:
            tab, nf, nt, mv = specaug
            K.cmvn_apply(raw, stats, lens, is_samples, feats_norm, x1, op_row_offset=2,
                         spec_masks=tab.to(device=dev, dtype=torch.int32).contiguous(), n_fmask=nf, n_tmask=nt,
                         mask_value=mv)
        else:
            K.cmvn_apply(raw, stats, lens, is_samples, feats_norm, x1, op_row_offset=2)
        seq_lens = self.buf("seq_lens", (B,), torch.int32)
        # subsampled lengths and the encoder padding mask [B, T] (fairseq lengths_to_padding_mask) in one launch
        self.pad_mask = torch.empty(B, sub_len(sub_len(m)), dtype=torch.bool, device=dev)
        K.seq_lens_mask(lens, is_samples, 2, seq_lens, self.pad_mask)
        return x1, m, seq_lens, feats_norm

    def subsample(self, x1: torch.Tensor, m: int, seq_lens: torch.Tensor) -> Tuple[torch.Tensor, int]:
        """Conv1d(k5,s2)+GLU twice as two GEMMs over strided windows -> fp32 residual stream [B*T, d]."""
        B, m_alloc, op = x1.shape[0], x1.shape[1], self.op_dtype
        c1, c2 = self.conv
        T1, mid = sub_len(m), c1["cout"] // 2
        T1_alloc = _even(T1 + 4)
        x2 = self.buf("x2", (B, T1_alloc, mid), op, zero=True)      # rows 0,1 and >= T1+2 stay zero
        K.gemm(a0=x1, a0_ld=2 * c1["cin"], a0_bs=m_alloc * c1["cin"], rows=T1, batches=B, w=c1["w"],
               n=c1["cout"], k=c1["k"] * c1["cin"], mode=K.EPI_GLU_OP, bias=c1["b"], out0=x2, out0_ld=mid,
               out0_bs=T1_alloc * mid, out_row_offset=2, block_n=self.block_n)
        T = sub_len(T1)
        x = self.buf("x", (B * T, self.d), torch.float32)
        K.gemm(a0=x2, a0_ld=2 * mid, a0_bs=T1_alloc * mid, rows=T, batches=B, w=c2["w"], n=c2["cout"],
               k=c2["k"] * mid, mode=K.EPI_GLU_POS_F32, bias=c2["b"], out0=x, out0_ld=self.d, out0_bs=T * self.d,
               scale=self.embed_scale, pos=self.pos_table(T + 2), seq_lens=seq_lens, block_n=self.block_n)
        return x, T

    def layer(self, L: dict, x: torch.Tensor, B: int, T: int, seq_lens: torch.Tensor, next_ln=None,
              h_f32: Optional[torch.Tensor] = None, h_out: Optional[torch.Tensor] = None) -> None:
        """One pre-LN encoder layer on the fp32 residual stream ``x`` (in place).

        ``next_ln`` = (gamma, beta) of the LayerNorm that FOLLOWS this layer (next layer's self_attn_layer_norm
        or the encoder's final layer_norm).  With d_model == 512 both LayerNorms of the layer are fused into
        the out_proj / fc2 GEMM epilogues (``mm_gemm_resid_ln``): on entry buffer ``h`` already holds
        LN1(x), on exit it (or ``h_out``) holds next_ln(x).  Otherwise the stand-alone LayerNorm kernel runs.
        """
        d, M, op, bn = self.d, B * T, self.op_dtype, self.block_n
        h = self.buf("h", (M, d), op)
        qkv = self.buf("qkv", (M, 3 * d), op)
        att = self.buf("att", (M, d), op)
        f = self.buf("ffn", (M, self.ffn), op)
        fused = self.fused_ln
        if not fused:
            K.layernorm(x, L["ln1_g"], L["ln1_b"], out_op=h)
        K.gemm(a0=h, a0_ld=d, rows=M, w=L["wqkv"], n=3 * d, k=d, mode=K.EPI_OP, bias=L["bqkv"], scale=64 ** -0.5,
               scale_cols=d, out0=qkv, out0_ld=3 * d, block_n=bn)
        K.self_attention(qkv, seq_lens, B, T, self.heads, att)
        if fused:
            K.gemm_resid_ln(att, L["wo"], L["bo"], x, L["ln2_g"], L["ln2_b"], h)
        else:
            K.gemm(a0=att, a0_ld=d, rows=M, w=L["wo"], n=d, k=d, mode=K.EPI_RESID_F32, bias=L["bo"], aux0=x,
                   aux_ld=d, out0=x, out0_ld=d, block_n=bn)
            K.layernorm(x, L["ln2_g"], L["ln2_b"], out_op=h)
        K.gemm(a0=h, a0_ld=d, rows=M, w=L["w1"], n=self.ffn, k=d, mode=K.EPI_RELU_OP, bias=L["b1"], out0=f,
               out0_ld=self.ffn, block_n=bn)
        if fused:
            K.gemm_resid_ln(f, L["w2"], L["b2"], x, next_ln[0], next_ln[1], h if h_out is None else h_out, h_f32)
        else:
            K.gemm(a0=f, a0_ld=self.ffn, rows=M, w=L["w2"], n=d, k=self.ffn, mode=K.EPI_RESID_F32, bias=L["b2"],
                   aux0=x, aux_ld=d, out0=x, out0_ld=d, block_n=bn)

    def fuse(self, j: int, text_f32: torch.Tensor, text_op: torch.Tensor, img: torch.Tensor,
             img_mask: Optional[torch.Tensor], B: int, T: int, out_tbc: torch.Tensor, img_dropout=None,
             attn_dropout=None, keep_scores: bool = False, tag: str = "") -> None:
        """fuse_img_feat for image type j; writes the fused states into out_tbc [T, B, d] fp32.
        img_dropout / attn_dropout = (p, seed, seed_dev, site): SA_image_dropout on the pre-normed image (:596) and
        SA_attention_dropout on the attention probabilities (fuse.py:111), training only.
        keep_scores: the training forward keeps S / P / K / V^T for the backward pass, i.e. runs the attention as
        scores GEMM -> softmax -> P V GEMM; otherwise (d % 256 == 0) the fused flash-style kernel runs and the scores
        never leave the SM.  tag: suffix of the per-call workspaces (several image types in one training step)."""
        enc, F, d, op, bn = self.enc, self.fusion[j], self.d, self.op_dtype, self.block_n
        M = B * T
        Tk_img, dk = img.shape[1], img.shape[2]
        if dk != F["dk"]:
            raise ValueError(f"image feature dim {dk} != image_feat_dim[{j}] = {F['dk']}")
        extra = 1 if F["bias_kv"] is not None else 0
        Tk = Tk_img + extra
        Tkp = _round_up(Tk, 8)
        img_op = self.buf(f"img_op{j}", (B * Tk_img, dk), op)
        if isinstance(img, StoredImages):   # features resident on the GPU in 16 bit: gather + pre-norm in one pass
            if self.img_ln is None:
                raise NotImplementedError("ImageFeatureStore batches need image_pre_norm: True (the gather is fused "
                                          "into the pre-norm kernel)")
            K.layernorm_gather(img.store.data, img.index, Tk_img, self.img_ln[0], self.img_ln[1], img_op)
        elif img.dtype in (torch.float16, torch.bfloat16) and self.img_ln is not None:
            # features shipped (or kept) in 16 bit: the pre-norm kernel reads them as they are, rows in order
            img = img.to(self.device, non_blocking=True).contiguous()
            K.layernorm_gather(img, None, Tk_img, self.img_ln[0], self.img_ln[1], img_op)
        else:
            img = img.to(self.device, non_blocking=True).float().contiguous()
            if self.img_ln is not None:
                K.layernorm(img.view(B * Tk_img, dk), self.img_ln[0], self.img_ln[1], out_op=img_op)
            else:
                K.convert(img.view(B * Tk_img, dk), img_op)
        if img_dropout is not None and img_dropout[0] > 0:
            K.dropout(img_op, img_op, img_dropout[0], img_dropout[1], img_dropout[3], seed_dev=img_dropout[2])
        mask = None
        if img_mask is not None:
            mask = img_mask.to(self.device).to(torch.uint8)
            if extra:
                mask = torch.nn.functional.pad(mask, (0, 1))
            mask = mask.contiguous()
        fused = self.fused_xattn and not keep_scores and not (attn_dropout is not None and attn_dropout[0] > 0)
        if fused:
            self._fuse_attention_fused(j, F, text_op, img_op, mask, B, T, Tk_img, Tk, dk)
            o = self.buf("o_img", (M, d), op)
        else:
            o = self._fuse_attention_unfused(j, F, text_op, img_op, mask, B, T, Tk_img, Tk, Tkp, dk, attn_dropout, tag)
        self._fuse_output(F, o, text_f32, text_op, B, T, out_tbc, tag)

    def _fuse_attention_fused(self, j, F, text_op, img_op, mask, B, T, Tk_img, Tk, dk) -> None:
        """K|V projection into ONE [B, Tk, 2d] tensor, q projection, then mm_cross_attention -> o_img."""
        d, op, bn, M = self.d, self.op_dtype, self.block_n, B * T
        extra = Tk - Tk_img
        new_kv = (f"kv{j}", (B, Tk, 2 * d), op) not in self._buf
        kv = self.buf(f"kv{j}", (B, Tk, 2 * d), op)
        if extra == 0:
            K.gemm(a0=img_op, a0_ld=dk, rows=B * Tk_img, w=F["wkv"], n=2 * d, k=dk, mode=K.EPI_OP, bias=F["bkv"],
                   out0=kv, out0_ld=2 * d, block_n=bn)
        else:
            K.gemm(a0=img_op, a0_ld=dk, a0_bs=Tk_img * dk, rows=Tk_img, batches=B, w=F["wkv"], n=2 * d, k=dk,
                   mode=K.EPI_OP, bias=F["bkv"], out0=kv, out0_ld=2 * d, out0_bs=Tk * 2 * d, block_n=bn)
            if new_kv:   # learned bias_k | bias_v = key / value row Tk_img; the GEMM never writes it (see below)
                kv[:, Tk_img, :d] = F["bias_kv"][0]
                kv[:, Tk_img, d:] = F["bias_kv"][1]
        q = self.buf("q_img", (M, d), op)
        K.gemm(a0=text_op, a0_ld=d, rows=M, w=F["wq"], n=d, k=d, mode=K.EPI_OP, bias=F["bq"], scale=d ** -0.5,
               scale_cols=d, out0=q, out0_ld=d, block_n=bn)
        o = self.buf("o_img", (M, d), op)
        kv2 = kv.view(B * Tk, 2 * d)
        K.cross_attention(q, T, kv2, 0, kv2, d, Tk, B, d, o, key_mask=mask)

    def _fuse_attention_unfused(self, j, F, text_op, img_op, mask, B, T, Tk_img, Tk, Tkp, dk, attn_dropout, tag=""):
        """scores GEMM -> softmax (+ attention dropout) -> P V GEMM with S, P, K and V^T kept in HBM (training).
        A key mask is burnt into S as -inf (``mm_mask_scores``), so the backward pass, which recomputes the
        probabilities from S, sees the same masked softmax."""
        d, op, bn, M = self.d, self.op_dtype, self.block_n, B * T
        extra = Tk - Tk_img
        new_kv = (f"k{j}", (B, Tk, d), op) not in self._buf
        kbuf = self.buf(f"k{j}", (B, Tk, d), op)
        vt = self.buf(f"vt_img{j}_{Tk}", (B, d, Tkp), op, zero=True)
        if extra == 0:   # keys are contiguous over the batch: one flat GEMM over all B * Tk image tokens
            K.gemm(a0=img_op, a0_ld=dk, rows=B * Tk_img, w=F["wkv"], n=2 * d, k=dk, mode=K.EPI_OP, bias=F["bkv"],
                   out0=kbuf, out0_ld=d, rows_per_seq=Tk_img, vt=vt, vt_col0=d, vt_rows=d, vt_ld=Tkp, block_n=bn)
        else:            # per-utterance key blocks of Tk_img + 1 rows: batched rows
            K.gemm(a0=img_op, a0_ld=dk, a0_bs=Tk_img * dk, rows=Tk_img, batches=B, w=F["wkv"], n=2 * d, k=dk,
                   mode=K.EPI_OP, bias=F["bkv"], out0=kbuf, out0_ld=d, out0_bs=Tk * d, vt=vt, vt_col0=d, vt_rows=d,
                   vt_ld=Tkp, block_n=bn)
        if extra and new_kv:
            # learned bias_k / bias_v are key / value number Tk_img (nn.MultiheadAttention add_bias_kv).  The K|V GEMM
            # only ever writes keys < Tk_img (its output map and the transposed-V store stop there), so the extra
            # key / value is written ONCE, when the workspace is created, not per forward.
            kbuf[:, Tk_img, :] = F["bias_kv"][0]
            vt[:, :, Tk_img] = F["bias_kv"][1]
        q = self.buf("q_img" + tag, (M, d), op)
        K.gemm(a0=text_op, a0_ld=d, rows=M, w=F["wq"], n=d, k=d, mode=K.EPI_OP, bias=F["bq"], scale=d ** -0.5,
               scale_cols=d, out0=q, out0_ld=d, block_n=bn)
        S = self.buf(f"S{j}", (B, T, Tkp), torch.float32)
        K.gemm(a0=q, a0_ld=d, a0_bs=T * d, rows=T, batches=B, w=kbuf, w_ld=d, w_bs=Tk * d, w_batched=True, n=Tk, k=d,
               mode=K.EPI_F32, out0=S, out0_ld=Tkp, out0_bs=T * Tkp, block_n=bn)
        P = self.buf(f"P{j}", (B, T, Tkp), op)
        if mask is not None:
            K.mask_scores(S, Tkp, M, Tk, mask, T)
        if attn_dropout is not None and attn_dropout[0] > 0:
            K.softmax_bwd(S, None, Tkp, M, T, Tk, None, Tkp, probs=P, drop_p=attn_dropout[0], seed=attn_dropout[1],
                          seed_dev=attn_dropout[2], site=attn_dropout[3])
        else:
            K.softmax_rows(S, Tkp, M, Tk, P, Tkp)
        o = self.buf("o_img" + tag, (M, d), op)
        K.gemm(a0=P, a0_ld=Tkp, a0_bs=T * Tkp, rows=T, batches=B, w=vt, w_ld=Tkp, w_bs=d * Tkp, w_batched=True, n=d,
               k=Tkp, mode=K.EPI_OP, out0=o, out0_ld=d, out0_bs=T * d, block_n=bn)
        return o

    def _fuse_output(self, F, o, text_f32, text_op, B, T, out_tbc, tag="") -> None:
        """out_proj of the attention, then the selective gate (or the plain residual) -> out_tbc [T, B, d]."""
        enc, d, op, bn, M = self.enc, self.d, self.op_dtype, self.block_n, B * T
        if enc.use_selective_gate:
            a_f32 = self.buf("attn_f32" + tag, (M, d), torch.float32)
            a_op = self.buf("attn_op" + tag, (M, d), op)
            K.gemm(a0=o, a0_ld=d, rows=M, w=F["wp"], n=d, k=d, mode=K.EPI_F32_OP, bias=F["bp"], out0=a_f32, out0_ld=d,
                   out1=a_op, out1_ld=d, block_n=bn)
            # rows batched per utterance so that a tile never straddles two utterances of the T x B x C output
            K.gemm(a0=a_op, a0_ld=d, a0_bs=T * d, a1=text_op, a1_ld=d, a1_bs=T * d, k_split=d, rows=T, batches=B,
                   w=F["wg"], n=d, k=2 * d, mode=K.EPI_GATE, bias=F["bg"], aux0=text_f32, aux1=a_f32, aux_ld=d,
                   out0=out_tbc, out0_ld=d, out_tbc=True, n_seqs=B, block_n=bn)
        else:
            K.gemm(a0=o, a0_ld=d, a0_bs=T * d, rows=T, batches=B, w=F["wp"], n=d, k=d, mode=K.EPI_RESID_F32,
                   bias=F["bp"], aux0=text_f32, aux_ld=d, out0=out_tbc, out0_ld=d, out_tbc=True, n_seqs=B, block_n=bn)

    # ------------------------------------------------------------------------------------------
    # full forward
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, src_tokens, src_lengths, imgs_list: List[torch.Tensor], img_masks_list: List,
                return_all_hiddens: bool = False, drop_audio: bool = False, drop_image: bool = False,
                training: bool = False):
        enc = self.enc
        if training and max(enc.dropout_p, getattr(enc, "SA_image_dropout", 0.0), getattr(enc, "SA_text_dropout", 0.0),
                            getattr(enc, "SA_attention_dropout", 0.0)) > 0:
            raise NotImplementedError(
                "EncoderEngine.forward is the inference forward (no dropout masks, no saved activations): a training-mode "
                "forward under torch.no_grad() with non-zero dropout is not defined here; the training step runs through "
                "TrainEngine.forward_train (enc.train() with gradients enabled), or call .eval()")
        x1, m, seq_lens, _ = self.frontend(src_tokens, src_lengths)
        B = x1.shape[0]
        x, T = self.subsample(x1, m, seq_lens)
        states = []
        text_f32 = self.buf("text_f32", (B * T, self.d), torch.float32)
        text_op = self.buf("text_op", (B * T, self.d), self.op_dtype)
        if self.fused_ln:   # LN1 of layer 0 is the only stand-alone LayerNorm; the rest ride on GEMM epilogues
            K.layernorm(x, self.layers[0]["ln1_g"], self.layers[0]["ln1_b"], out_op=self.buf("h", (B * T, self.d),
                                                                                              self.op_dtype))
        for i, L in enumerate(self.layers):
            last = i + 1 == len(self.layers)
            nxt = (self.ln_g, self.ln_b) if last else (self.layers[i + 1]["ln1_g"], self.layers[i + 1]["ln1_b"])
            self.layer(L, x, B, T, seq_lens, next_ln=nxt, h_f32=text_f32 if last else None,
                       h_out=text_op if last else None)
            if return_all_hiddens:
                states.append(x.view(B, T, self.d).transpose(0, 1).contiguous())
        if not self.fused_ln:
            K.layernorm(x, self.ln_g, self.ln_b, out_op=text_op, out_f32=text_f32)
        mask = self.pad_mask            # written by frontend()
        if imgs_list and self.fusion:
            if drop_audio:
                text_f32.zero_()
                text_op.zero_()
            outs = []
            for j, (img, img_mask) in enumerate(zip(imgs_list, img_masks_list)):
                if drop_image:   # modality dropout: every image tensor zeroed (reference :504-505)
                    img = torch.zeros(tuple(img.shape), dtype=torch.float32, device=self.device)
                res = torch.empty(T, B, self.d, dtype=torch.float32, device=self.device)
                self.fuse(j, text_f32, text_op, img, img_mask, B, T, res)
                outs.append(res)
            out = outs[0]
            for o in outs[1:]:
                out = out + o
        else:
            out = text_f32.view(B, T, self.d).transpose(0, 1).contiguous()
        return {
            "encoder_out": [out],
            "encoder_padding_mask": [mask],
            "encoder_embedding": [],
            "encoder_states": states,
            "src_tokens": [],
            "src_lengths": [],
        }
