"""CPU emulation of the C-ABI kernel wrappers (``mm_s2ut_b200.kernels``) -- TEST INFRASTRUCTURE ONLY.

The engines (``engine.py`` / ``training.py``) are host logic: they choose buffers, strides, batch offsets and the
order of launches.  This module re-states every wrapper they call as plain PyTorch on CPU tensors, honouring the same
pointer / leading-dimension / batch-stride arguments (``as_strided`` over the tensor's storage), so that the host
logic -- in particular the whole backward orchestration -- can be checked against autograd over the oracle WITHOUT a
GPU (``tests/test_host_training.py``).  It is never imported by the package; the product path has no CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np
import torch

EPI_OP, EPI_RELU_OP, EPI_RESID_F32, EPI_GLU_OP, EPI_GLU_POS_F32, EPI_F32_OP, EPI_GATE, EPI_F32, EPI_MASK_OP = range(9)
launch_count = 0
timing = None
scope = ""


def _v(t, sizes, strides, extra_offset=0):
    """Strided view over t's storage starting at t's first element (+ extra_offset elements)."""
    return t.as_strided(tuple(int(s) for s in sizes), tuple(int(s) for s in strides), t.storage_offset() + extra_offset)


def dtype_code(dt):
    return {torch.bfloat16: 0, torch.float16: 1}[dt]


def fbank_tables(device):
    return torch.zeros(1)


def _frames(n):
    return 0 if n < 400 else 1 + (n - 400) // 160


def fbank(wav, n_samples, feats, tables):
    from oracle import fbank as ofb

    for b in range(wav.shape[0]):
        n = int(n_samples[b])
        f = ofb.kaldi_fbank_ta(wav[b, :n].float().numpy())
        feats[b, : f.shape[0]] = torch.from_numpy(f)


def _lens_frames(lens, is_samples):
    return [(_frames(int(v)) if is_samples else int(v)) for v in lens]


def cmvn_stats(feats, lens, lengths_are_samples, mean_std):
    for b, m in enumerate(_lens_frames(lens, lengths_are_samples)):
        x = feats[b, :m].numpy()
        mean = x.mean(0)
        var = (x ** 2).sum(0) / m - mean ** 2
        mean_std.view(-1, 2, 80)[b, 0] = torch.from_numpy(mean)
        mean_std.view(-1, 2, 80)[b, 1] = torch.from_numpy(np.sqrt(np.maximum(var, 1e-10)))


def cmvn_apply(feats, stats, lens, lengths_are_samples, out_f32, out_op, op_row_offset=0, spec_masks=None, n_fmask=0,
               n_tmask=0, mask_value=0.0):
    if out_op is not None:
        out_op.zero_()
    if out_f32 is not None:
        out_f32.zero_()
    for b, m in enumerate(_lens_frames(lens, lengths_are_samples)):
        x = feats[b, :m]
        if stats is not None:
            x = (x - stats.view(-1, 2, 80)[b, 0]) / stats.view(-1, 2, 80)[b, 1]
        if spec_masks is not None:
            x = x.clone()
            row = spec_masks.view(len(lens), -1)[b].tolist()
            for i in range(n_fmask):
                x[:, row[2 * i]:row[2 * i] + row[2 * i + 1]] = mask_value
            for i in range(n_fmask, n_fmask + n_tmask):
                x[row[2 * i]:row[2 * i] + row[2 * i + 1]] = mask_value
        if out_f32 is not None:
            out_f32[b, :m] = x
        if out_op is not None:
            out_op[b, op_row_offset:op_row_offset + m] = x.to(out_op.dtype)


def seq_lens(lens, lengths_are_samples, n_layers, out):
    for b, m in enumerate(_lens_frames(lens, lengths_are_samples)):
        for _ in range(n_layers):
            m = (m - 1) // 2 + 1
        out[b] = m


def seq_lens_mask(lens, lengths_are_samples, n_layers, out, mask):
    seq_lens(lens, lengths_are_samples, n_layers, out)
    padding_mask(out, mask.shape[1], mask)


def padding_mask(seq_lens_, T, out):
    out.copy_(torch.arange(T)[None, :] >= seq_lens_[:, None])


def _operand(t, mn, kbatch, hm, rows, k, ld, bs, batches, heads, hs, k_total):
    """[batches, rows, k] fp32 view of a GEMM operand under the layout flags of mm_gemm_args."""
    if mn and kbatch:
        kt = k_total if k_total > 0 else batches * k
        full = torch.zeros(batches * k, rows)
        take = min(kt, batches * k)
        full[:take] = _v(t, (take, rows), (ld, 1)).float()
        return full.view(batches, k, rows).transpose(1, 2)
    if hm:
        nseq = batches // heads
        if mn:
            x = _v(t, (nseq, heads, k, rows), (bs, hs, ld, 1))
            return x.reshape(batches, k, rows).float().transpose(1, 2)
        return _v(t, (nseq, heads, rows, k), (bs, hs, ld, 1)).reshape(batches, rows, k).float()
    if mn:
        return _v(t, (batches, k, rows), (bs, ld, 1)).float().transpose(1, 2)
    return _v(t, (batches, rows, k), (bs, ld, 1)).float()


def gemm(*, a0, w, rows, n, k, mode, out0, a0_ld, out0_ld, batches=1, a0_bs=0, a1=None, a1_ld=0, a1_bs=0, k_split=0,
         w_ld=None, w_bs=0, w_batched=False, bias=None, scale=1.0, scale_cols=0, out0_bs=0, out1=None, out1_ld=0,
         out1_bs=0, aux0=None, aux1=None, aux_ld=0, rows_per_seq=0, out_tbc=False, n_seqs=0, out_row_offset=0, vt=None,
         vt_col0=0, vt_rows=0, vt_ld=0, pos=None, seq_lens=None, block_n=0, a_mn=False, w_mn=False, a_kbatch=False,
         w_kbatch=False, a_hm=False, w_hm=False, out_hm=False, heads=0, head_stride=0, a_k_total=0, w_k_total=0,
         drop=None):
    global launch_count
    launch_count += 1
    w_ld = k if w_ld is None else w_ld
    k0 = k_split if a1 is not None else k
    if a_kbatch or w_kbatch:
        assert k % 64 == 0
    A = _operand(a0, a_mn, a_kbatch, a_hm, rows, k0, a0_ld, a0_bs, batches, heads, head_stride, a_k_total)
    if a1 is not None:
        A = torch.cat([A, _v(a1, (batches, rows, k - k0), (a1_bs, a1_ld, 1)).float()], -1)
    if w_batched:
        W = _operand(w, w_mn, w_kbatch, w_hm, n, k, w_ld, w_bs, batches, heads, head_stride, w_k_total)
    else:
        W = _operand(w, w_mn, False, False, n, k, w_ld, 0, 1, 0, 0, 0)
    acc = A @ W.transpose(-1, -2)
    if bias is not None:
        acc = acc + bias[:n]
    op = w.dtype

    def out_view(t, cols, ld, bs, rows_=rows, roff=0):
        if out_hm:
            assert mode == EPI_OP and n % 64 == 0
            return _HeadView(t, batches, heads, rows_, cols, bs, ld, head_stride)
        if out_tbc:
            return _v(t, (batches, rows_, cols), (ld, n_seqs * ld, 1))
        return _v(t, (batches, rows_, cols), (bs, ld, 1), roff * ld)

    if mode in (EPI_OP, EPI_RELU_OP):
        if mode == EPI_OP and scale_cols > 0:
            acc[..., :scale_cols] *= scale
        if mode == EPI_RELU_OP:
            acc = acc.relu()
            if drop is not None and drop[0] > 0:
                acc = acc * _keep_scale(acc.numel(), drop[0], drop[1], drop[3], drop[2]).view(acc.shape)
        n_out = vt_col0 if (mode == EPI_OP and vt is not None) else n
        if n_out > 0:
            out_view(out0, n_out, out0_ld, out0_bs, roff=out_row_offset).copy_(acc[..., :n_out].to(op))
        if mode == EPI_OP and vt is not None:
            tail = acc[..., vt_col0:].to(op)                                  # [batches, rows, n - vt_col0]
            if rows_per_seq > 0:
                tail = tail.reshape(rows // rows_per_seq, rows_per_seq, -1)
            V = _v(vt, (tail.shape[0], tail.shape[2], tail.shape[1]), (vt_rows * vt_ld, vt_ld, 1))
            V.copy_(tail.transpose(1, 2))
    elif mode in (EPI_GLU_OP, EPI_GLU_POS_F32):
        t = acc.reshape(batches, rows, n // 256, 2, 128)
        y = (t[..., 0, :] * torch.sigmoid(t[..., 1, :])).reshape(batches, rows, n // 2)
        if mode == EPI_GLU_OP:
            out_view(out0, n // 2, out0_ld, out0_bs, roff=out_row_offset).copy_(y.to(op))
        else:
            y = y * scale
            for b in range(batches):
                L = min(int(seq_lens[b]), rows) if seq_lens is not None else rows
                y[b, :L] += pos[2:2 + L, : n // 2]
            out_view(out0, n // 2, out0_ld, out0_bs).copy_(y)
    elif mode == EPI_F32:
        out_view(out0, n, out0_ld, out0_bs).copy_(acc)
    elif mode == EPI_F32_OP:
        out_view(out0, n, out0_ld, out0_bs).copy_(acc)
        _v(out1, (batches, rows, n), (out1_bs, out1_ld, 1)).copy_(acc.to(op))
    elif mode == EPI_RESID_F32:
        x = _v(aux0, (batches, rows, n), (rows * aux_ld, aux_ld, 1)).clone()
        out_view(out0, n, out0_ld, out0_bs).copy_(acc + x)
    elif mode == EPI_GATE:
        x = _v(aux0, (batches, rows, n), (rows * aux_ld, aux_ld, 1))
        o = _v(aux1, (batches, rows, n), (rows * aux_ld, aux_ld, 1))
        g = torch.sigmoid(acc)
        out_view(out0, n, out0_ld, out0_bs).copy_((1 - g) * x + g * o)
    elif mode == EPI_MASK_OP:
        m = _v(aux0, (batches, rows, n), (rows * aux_ld, aux_ld, 1)).float() > 0
        out_view(out0, n, out0_ld, out0_bs).copy_(torch.where(m, acc * scale, torch.zeros_like(acc)).to(op))
    else:
        raise ValueError(mode)


class _HeadView:
    """copy_ target for head-mode outputs: [batches = (seq, head), rows, cols] -> t[seq][row][head * hs + col]."""

    def __init__(self, t, batches, heads, rows, cols, bs, ld, hs):
        self.v = _v(t, (batches // heads, heads, rows, cols), (bs, hs, ld, 1))
        self.shape = (batches // heads, heads, rows, cols)

    def copy_(self, src):
        self.v.copy_(src.reshape(self.shape))


def gemm_resid_ln(a, w, bias, x, gamma, beta, h_op, h_f32=None, eps=1e-5, x_out=None, drop=None):
    global launch_count
    launch_count += 1
    x_out = x if x_out is None else x_out
    y = a.float() @ w.float().t() + bias
    if drop is not None and drop[0] > 0:
        y = y * _keep_scale(y.numel(), drop[0], drop[1], drop[3], drop[2]).view(y.shape)
    x_out.copy_(x + y)
    y = torch.nn.functional.layer_norm(x_out, (x_out.shape[-1],), gamma, beta, eps)
    h_op.copy_(y.to(h_op.dtype))
    if h_f32 is not None:
        h_f32.copy_(y)


def layernorm(x, gamma, beta, out_op=None, out_f32=None, eps=1e-5):
    dim = x.shape[-1]
    y = torch.nn.functional.layer_norm(x.reshape(-1, dim), (dim,), gamma, beta, eps)
    if out_op is not None:
        out_op.view(-1, dim).copy_(y.to(out_op.dtype))
    if out_f32 is not None:
        out_f32.view(-1, dim).copy_(y)


def self_attention_drop_supported(seq):
    return 128 < seq <= 256


def _attn_keep(batch, heads, seq, drop, kv=None):
    """keep / (1 - p) over [batch*heads][Lp][Tp] (Lp / Tp = seq / kv rounded up to 64), cut to [batch, heads, seq, kv]."""
    kv = seq if kv is None else kv
    Lp, Tp = (seq + 63) // 64 * 64, (kv + 63) // 64 * 64
    m = _keep_scale(batch * heads * Lp * Tp, drop[0], drop[1], drop[3], drop[2]).view(batch, heads, Lp, Tp)
    return m[:, :, :seq, :kv]


def self_attention(qkv, seq_lens_, batch, seq, heads, out, lse=None, drop=None):
    d = heads * 64
    x = qkv.float().view(batch, seq, 3, heads, 64)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    s = q @ k.transpose(-1, -2)
    mask = torch.arange(seq)[None, :] >= seq_lens_[:, None]
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    if lse is not None:
        lse.copy_(torch.logsumexp(s, -1).reshape(lse.shape))
    p = s.softmax(-1)
    if drop is not None and drop[0] > 0:
        p = p * _attn_keep(batch, heads, seq, drop)
    p = p.to(qkv.dtype).float()
    o = (p @ v).permute(0, 2, 1, 3).reshape(batch * seq, d)
    out.copy_(o.to(out.dtype))


def cross_attention(q, q_len, k, k_col0, v, v_col0, kv_len, batch, d_model, out, key_mask=None, lse=None,
                    kv_batch_stride=0):
    Q = q[:, :d_model].float().view(batch, q_len, d_model)
    Kx = _v(k, (batch, kv_len, d_model), (kv_batch_stride or kv_len * k.stride(0), k.stride(0), 1), k_col0).float()
    V = _v(v, (batch, kv_len, d_model), (kv_batch_stride or kv_len * v.stride(0), v.stride(0), 1), v_col0).float()
    s = Q @ Kx.transpose(-1, -2)
    if key_mask is not None:
        s = s.masked_fill(key_mask[:, None, :kv_len].bool(), float("-inf"))
    if lse is not None:
        lse.copy_(torch.logsumexp(s, -1).reshape(-1))
    p = s.softmax(-1).to(q.dtype).float()
    out.copy_((p @ V).reshape(batch * q_len, d_model).to(out.dtype))


def mask_scores(scores, ld, rows, n_keys, key_mask, rows_per_seq):
    s = _v(scores, (rows, n_keys), (ld, 1))
    s.masked_fill_(key_mask.bool().repeat_interleave(rows_per_seq, 0)[:, :n_keys], float("-inf"))


def softmax_rows(scores, ld_in, rows, n_keys, probs, ld_out, key_mask=None, rows_per_seq=0):
    s = _v(scores, (rows, n_keys), (ld_in, 1))
    if key_mask is not None:
        s = s.masked_fill(key_mask.bool().repeat_interleave(rows_per_seq, 0)[:, :n_keys], float("-inf"))
    P = _v(probs, (rows, ld_out), (ld_out, 1))
    P.zero_()
    P[:, :n_keys] = s.softmax(-1).to(probs.dtype)


def _qkv_heads(q, q_col0, q_len, k, k_col0, v, v_col0, kv_len, batch, heads):
    d = heads * 64
    Q = q[:, q_col0:q_col0 + d].float().view(batch, q_len, heads, 64).permute(0, 2, 1, 3)
    Kx = k[:, k_col0:k_col0 + d].float().view(batch, kv_len, heads, 64).permute(0, 2, 1, 3)
    V = v[:, v_col0:v_col0 + d].float().view(batch, kv_len, heads, 64).permute(0, 2, 1, 3)
    return Q, Kx, V


def attention_bwd_scores(q, q_col0, q_len, k, k_col0, v, v_col0, kv_len, kv_lens, batch, heads, dout, out, lse, probs,
                         dscores, causal=False):
    Q, Kx, V = _qkv_heads(q, q_col0, q_len, k, k_col0, v, v_col0, kv_len, batch, heads)
    s = Q @ Kx.transpose(-1, -2)
    dead = torch.zeros(batch, 1, q_len, kv_len, dtype=torch.bool)
    if kv_lens is not None:
        dead = dead | (torch.arange(kv_len)[None, :] >= kv_lens[:, None])[:, None, None, :]
    if causal:
        dead = dead | (torch.arange(kv_len)[None, :] > torch.arange(q_len)[:, None])[None, None]
    P = torch.exp(s - lse.view(batch, heads, q_len, 1)).masked_fill(dead, 0.0)
    dO = dout[:, :heads * 64].float().view(batch, q_len, heads, 64).permute(0, 2, 1, 3)
    O = out[:, :heads * 64].float().view(batch, q_len, heads, 64).permute(0, 2, 1, 3)
    delta = (dO * O).sum(-1, keepdim=True)
    dS = P * (dO @ V.transpose(-1, -2) - delta)
    probs.view(batch, heads, probs.shape[1], probs.shape[2])[:, :, :q_len, :kv_len] = P.to(probs.dtype)
    dscores.view(batch, heads, probs.shape[1], probs.shape[2])[:, :, :q_len, :kv_len] = dS.to(dscores.dtype)


def attention_bwd_general_scratch_floats(kv_len):
    return 1


def attention_bwd_general(q, q_len, k, v, kv_len, kv_lens, batch, heads, dout, out, lse, dq, dk, dv, scratch, causal=False,
                          drop=None):
    global launch_count
    launch_count += 1
    d = heads * 64
    Q, Kx, V = _qkv_heads(q, 0, q_len, k, 0, v, 0, kv_len, batch, heads)
    s = Q @ Kx.transpose(-1, -2)
    dead = torch.zeros(batch, 1, q_len, kv_len, dtype=torch.bool)
    if kv_lens is not None:
        dead = dead | (torch.arange(kv_len)[None, :] >= kv_lens[:, None])[:, None, None, :]
    if causal:
        dead = dead | (torch.arange(kv_len)[None, :] > torch.arange(q_len)[:, None])[None, None]
    P = torch.exp(s - lse.view(batch, heads, q_len, 1)).masked_fill(dead, 0.0)
    dO = dout[:, :d].float().view(batch, q_len, heads, 64).permute(0, 2, 1, 3)
    O = out[:, :d].float().view(batch, q_len, heads, 64).permute(0, 2, 1, 3)
    mk = _attn_keep(batch, heads, q_len, drop, kv_len) if (drop is not None and drop[0] > 0) else 1.0
    dS = P * ((dO @ V.transpose(-1, -2)) * mk - (dO * O).sum(-1, keepdim=True))
    P = P * mk
    P, dS = P.to(q.dtype).float(), dS.to(q.dtype).float()
    back = lambda x, L: x.permute(0, 2, 1, 3).reshape(batch * L, d)
    dq[:, :d] = back(dS @ Kx * 0.125, q_len).to(dq.dtype)
    dk[:, :d] = back(dS.transpose(-1, -2) @ Q, kv_len).to(dk.dtype)
    dv[:, :d] = back(P.transpose(-1, -2) @ dO, kv_len).to(dv.dtype)


def attention_bwd_fused(qkv, seq_len, kv_lens, batch, heads, dout, out, lse, dqkv, drop=None):
    global launch_count
    launch_count += 1
    d = heads * 64
    Q, Kx, V = _qkv_heads(qkv, 0, seq_len, qkv, d, qkv, 2 * d, seq_len, batch, heads)
    s = Q @ Kx.transpose(-1, -2)
    dead = torch.zeros(batch, 1, seq_len, seq_len, dtype=torch.bool)
    if kv_lens is not None:
        dead = dead | (torch.arange(seq_len)[None, :] >= kv_lens[:, None])[:, None, None, :]
    P = torch.exp(s - lse.view(batch, heads, seq_len, 1)).masked_fill(dead, 0.0)
    dO = dout[:, :d].float().view(batch, seq_len, heads, 64).permute(0, 2, 1, 3)
    O = out[:, :d].float().view(batch, seq_len, heads, 64).permute(0, 2, 1, 3)
    mk = _attn_keep(batch, heads, seq_len, drop) if (drop is not None and drop[0] > 0) else 1.0
    dS = P * ((dO @ V.transpose(-1, -2)) * mk - (dO * O).sum(-1, keepdim=True))
    P = P * mk
    P, dS = P.to(qkv.dtype).float(), dS.to(qkv.dtype).float()
    back = lambda x: x.permute(0, 2, 1, 3).reshape(batch * seq_len, d).to(dqkv.dtype)
    dqkv[:, :d] = back(dS @ Kx * 0.125)
    dqkv[:, d:2 * d] = back(dS.transpose(-1, -2) @ Q)
    dqkv[:, 2 * d:3 * d] = back(P.transpose(-1, -2) @ dO)


def attention(q, q_col0, q_len, k, k_col0, v, v_col0, kv_len, kv_lens, batch, heads, out, causal=False, lse=None,
              drop=None):
    d = heads * 64
    Q = q[:, q_col0:q_col0 + d].float().view(batch, q_len, heads, 64).permute(0, 2, 1, 3)
    Kx = k[:, k_col0:k_col0 + d].float().view(batch, kv_len, heads, 64).permute(0, 2, 1, 3)
    V = v[:, v_col0:v_col0 + d].float().view(batch, kv_len, heads, 64).permute(0, 2, 1, 3)
    s = Q @ Kx.transpose(-1, -2)
    if kv_lens is not None:
        s = s.masked_fill((torch.arange(kv_len)[None, :] >= kv_lens[:, None])[:, None, None, :], float("-inf"))
    if causal:
        s = s.masked_fill(torch.arange(kv_len)[None, :] > torch.arange(q_len)[:, None], float("-inf"))
    if lse is not None:
        lse.copy_(torch.logsumexp(s, -1).reshape(lse.shape))
    p = s.softmax(-1)
    if drop is not None and drop[0] > 0:
        p = p * _attn_keep(batch, heads, q_len, drop, kv_len)
    p = p.to(q.dtype).float()
    out.copy_((p @ V).permute(0, 2, 1, 3).reshape(batch * q_len, d).to(out.dtype))


def embed_tokens(tokens, padding_idx, table, scale, pos_table, out):
    B, L = tokens.shape
    ne = tokens.ne(padding_idx)
    pos = torch.cumsum(ne, 1) * ne + padding_idx
    out.view(B, L, -1).copy_(scale * table[tokens] + pos_table[pos])


def label_smoothed_nll(logits, vocab, target, padding_idx, epsilon):
    lp = torch.log_softmax(logits[:, :vocab], -1)
    t = target.view(-1, 1)
    pad = t.eq(padding_idx)
    nll = (-lp.gather(1, t)).masked_fill(pad, 0).sum()
    smooth = (-lp.sum(1, keepdim=True)).masked_fill(pad, 0).sum()
    eps_i = epsilon / (vocab - 1)
    return (1 - epsilon - eps_i) * nll + eps_i * smooth, nll


def label_smoothed_nll_bwd(logits, vocab, target, padding_idx, epsilon, dlogits, grad_scale=1.0):
    p = torch.softmax(logits[:, :vocab], -1)
    t = target.view(-1)
    eps_i = epsilon / (vocab - 1)
    onehot = torch.nn.functional.one_hot(t, vocab).float()
    g = (1 - epsilon - eps_i) * (p - onehot) + eps_i * (vocab * p - 1)
    g = g * t.ne(padding_idx)[:, None] * grad_scale
    dlogits.zero_()
    dlogits[:, :vocab] = g.to(dlogits.dtype)


def embed_tokens_bwd(tokens, padding_idx, dx, scale, table_grad):
    t = tokens.view(-1)
    keep = t.ne(padding_idx)
    table_grad.view(-1, dx.shape[-1]).index_add_(0, t[keep], scale * dx.view(t.numel(), -1)[keep])


def convert(x, out):
    out.view(-1).copy_(x.reshape(-1).to(out.dtype))


# ---- training-step variant --------------------------------------------------------------------------------
def pack_t(x, *, rows, cols, in_ld, out_n=None, n_ld=0, out_t=None, t_ld=0, t_cols_pad=0, batches=1, nb1=1, in_bs0=0,
           in_bs1=0, n_bs0=0, n_bs1=0, t_bs0=0, t_bs1=0, mask=None, mask_ld=0, scale=1.0):
    global launch_count
    launch_count += 1
    assert cols % 2 == 0 and all(s % 2 == 0 for s in (in_ld, in_bs0, in_bs1, n_ld, t_ld, n_bs0, n_bs1, t_bs0, t_bs1))
    op = (out_n if out_n is not None else out_t).dtype
    rows_even = rows + (rows & 1)
    if out_t is not None:
        t_cols_pad = max(t_cols_pad, rows_even)
        assert t_cols_pad % 2 == 0 and t_ld >= t_cols_pad
    for z in range(batches):
        b0, b1 = divmod(z, nb1)
        v = _v(x, (rows, cols), (in_ld, 1), b0 * in_bs0 + b1 * in_bs1).float() * scale
        if mask is not None:
            v = v * (_v(mask, (rows, cols), (mask_ld, 1)).float() > 0)
        v = v.to(op)
        if out_n is not None:
            _v(out_n, (rows, cols), (n_ld, 1), b0 * n_bs0 + b1 * n_bs1).copy_(v)
        if out_t is not None:
            T = _v(out_t, (cols, t_cols_pad), (t_ld, 1), b0 * t_bs0 + b1 * t_bs1)
            T[:, :rows] = v.t()
            T[:, rows:] = 0


def rowsum(x, ld, rows, cols, out, accumulate=False):
    s = _v(x, (rows, cols), (ld, 1)).float().sum(1)
    out[:rows] = out[:rows] + s if accumulate else s


def colsum_blocks(rows):
    return (rows + 127) // 128


def colsum(x, ld, rows, cols, partials, period=0, valid=0):
    nb = colsum_blocks(rows)
    v = _v(x, (rows, cols), (ld, 1)).float()
    if period > 0:
        v = v * ((torch.arange(rows) % period) < valid)[:, None]
    p = partials[: nb * cols].view(nb, cols)
    p.zero_()
    p[0] = v.sum(0)
    return nb


def heads_gemm(a, a_ld, a_bs, transposed, w, w_ld, w_bs, out, out_ld, out_bs, rows, k, batch, heads, scale=1.0):
    global launch_count
    launch_count += 1
    BH = batch * heads
    A = (_v(a, (BH, k, rows), (a_bs, a_ld, 1)).transpose(1, 2) if transposed else _v(a, (BH, rows, k), (a_bs, a_ld, 1)))
    W = _v(w, (batch, heads, k, 64), (w_bs, 64, w_ld, 1)).float().reshape(BH, k, 64)
    o = (A.float() @ W * scale).view(batch, heads, rows, 64)
    _v(out, (batch, heads, rows, 64), (out_bs, 64, out_ld, 1)).copy_(o.to(out.dtype))


def wgrad_grouped(groups, tokens, accumulate=False):
    global launch_count
    launch_count += 1
    tokens_all = tokens
    for grp in groups:
        dy, dy_ld, x, x_ld, out, out_ld, n_out, k_in, bias = grp[:9]
        tokens = (grp[9] if len(grp) > 9 and grp[9] else tokens_all)
        dyf = _v(dy, (tokens, n_out), (dy_ld, 1)).float()
        if k_in > 0:
            g = dyf.t() @ _v(x, (tokens, k_in), (x_ld, 1)).float()
            o = _v(out, (n_out, k_in), (out_ld, 1))
            o.copy_(o + g if accumulate else g)
        if bias is not None:
            b = bias.view(-1)[:n_out]
            b.copy_(b + dyf.sum(0) if accumulate else dyf.sum(0))


def reduce_partials(part, n_partials, stride, n, out, accumulate=False, part_offset=0):
    s = _v(part, (n_partials, n), (stride, 1), part_offset).sum(0)
    out.view(-1)[:n] = out.view(-1)[:n] + s if accumulate else s


def reduce_partials_many(jobs):
    for part, S, stride, n, out, acc in jobs:
        reduce_partials(part, S, stride, n, out, acc)


def layernorm_bwd_blocks():
    return 4


def layernorm_bwd(x, gamma, dy, partials, dx=None, resid=None, eps=1e-5, dx_op=None, drop=None):
    dim = x.shape[-1]
    xr = x.reshape(-1, dim)
    dyr = dy.reshape(-1, dim)
    mean = xr.mean(1, keepdim=True)
    rstd = torch.rsqrt(((xr - mean) ** 2).mean(1, keepdim=True) + eps)
    xhat = (xr - mean) * rstd
    dyg = dyr * gamma
    d = rstd * (dyg - dyg.mean(1, keepdim=True) - xhat * (dyg * xhat).mean(1, keepdim=True))
    if dx is not None:
        dx.view(-1, dim).copy_(d + (resid.reshape(-1, dim) if resid is not None else 0))
        if dx_op is not None:
            g16 = dx.view(-1, dim)
            if drop is not None and drop[0] > 0:
                g16 = g16 * _keep_scale(g16.numel(), drop[0], drop[1], drop[3], drop[2]).view(g16.shape)
            dx_op.view(-1, dim).copy_(g16.to(dx_op.dtype))
    p = partials[: layernorm_bwd_blocks() * 2 * dim].view(layernorm_bwd_blocks(), 2, dim)
    p.zero_()
    p[0, 0] = (dyr * xhat).sum(0)
    p[0, 1] = dyr.sum(0)


def gemm_ln_bwd_partial_rows(rows):
    return 4


def gemm_ln_bwd(dy, w, x, gamma, g, g_op, partials, eps=1e-5, drop=None):
    global launch_count
    launch_count += 1
    rows = dy.shape[0]
    dh = dy.float() @ w.float()
    xr = x.view(rows, 512)
    mean = xr.mean(1, keepdim=True)
    rstd = torch.rsqrt(((xr - mean) ** 2).mean(1, keepdim=True) + eps)
    xhat = (xr - mean) * rstd
    dyg = dh * gamma
    gv = g.view(rows, 512)
    gv.add_(rstd * (dyg - dyg.mean(1, keepdim=True) - xhat * (dyg * xhat).mean(1, keepdim=True)))
    g16 = gv
    if drop is not None and drop[0] > 0:
        g16 = gv * _keep_scale(gv.numel(), drop[0], drop[1], drop[3], drop[2]).view(gv.shape)
    g_op.view(rows, 512).copy_(g16.to(g_op.dtype))
    p = partials[: 4 * 1024].view(4, 2, 512)
    p.zero_()
    p[0, 0] = (dh * xhat).sum(0)
    p[0, 1] = dh.sum(0)
    return 4


def softmax_bwd(scores, dprobs, ld_in, rows, rows_per_batch, n_keys, dscores, ld_out, probs=None, kv_lens=None, heads=1,
                valid_rows=0, causal=False, ld_dprobs=None, drop_p=0.0, seed=0, seed_dev=None, site=0):
    S = _v(scores, (rows, n_keys), (ld_in, 1))
    D = None if dprobs is None else _v(dprobs, (rows, n_keys), (ld_in if ld_dprobs is None else ld_dprobs, 1)).float()
    G = None if dscores is None else _v(dscores, (rows, ld_out), (ld_out, 1))
    Pv = None if probs is None else _v(probs, (rows, ld_out), (ld_out, 1))
    mult = torch.ones(rows, ld_out)
    if drop_p > 0:
        if seed_dev is not None:
            seed = int(seed) + int(seed_dev.view(-1)[0])
        g = torch.Generator().manual_seed((int(seed) * 1000003 + int(site)) % (2 ** 63 - 1))
        mult = ((torch.rand(rows * ld_out, generator=g) >= drop_p).float() / (1.0 - drop_p)).view(rows, ld_out)
    vr = valid_rows if valid_rows > 0 else rows_per_batch
    for r0 in range(0, rows, rows_per_batch):
        valid = n_keys if kv_lens is None else min(n_keys, int(kv_lens[(r0 // rows_per_batch) // heads]))
        sl = slice(r0, min(rows, r0 + vr))
        if G is not None:
            G[sl] = 0
        if Pv is not None:
            Pv[sl] = 0
        sc = S[sl, :valid].clone()
        if causal:
            nq = sc.shape[0]
            sc = sc.masked_fill(torch.arange(valid)[None, :] > torch.arange(nq)[:, None], float("-inf"))
        p = sc.softmax(-1)
        m = mult[sl, :valid]
        if G is not None:
            d = D[sl, :valid] * m
            G[sl, :valid] = (p * (d - (p * d).sum(1, keepdim=True))).to(dscores.dtype)
        if Pv is not None:
            Pv[sl, :valid] = (p * m).to(probs.dtype)


def glu_bwd(pre, dy, rows, n, dpre, scale=1.0):
    a, b = pre.view(rows, 2 * n)[:, :n], pre.view(rows, 2 * n)[:, n:]
    g = dy.view(rows, n) * scale
    s = torch.sigmoid(b)
    dpre.view(rows, 2 * n)[:, :n] = (g * s).to(dpre.dtype)
    dpre.view(rows, 2 * n)[:, n:] = (g * a * s * (1 - s)).to(dpre.dtype)


def gate_bwd(z, dres_tbc, text, attn, B, T, d, dz, dcat):
    dres = dres_tbc.view(T, B, d).transpose(0, 1).reshape(B * T, d)
    g = torch.sigmoid(z.view(B * T, d))
    dz.view(B * T, d).copy_((dres * (attn.view(B * T, d) - text.view(B * T, d)) * g * (1 - g)).to(dz.dtype))
    dcat.view(B * T, 2 * d)[:, :d] = dres * g
    dcat.view(B * T, 2 * d)[:, d:] = dres * (1 - g)


def tbc_to_btc(x_tbc, B, T, d, out):
    out.view(B, T, d).copy_(x_tbc.view(T, B, d).transpose(0, 1))


def col2im_k5s2(dcol, B, t_out, t_in, C_, dx):
    dc = dcol.view(B, t_out, 5, C_)
    o = dx.view(B, t_in, C_)
    o.zero_()
    for t in range(t_out):
        for k in range(5):
            s = 2 * t + k - 2
            if 0 <= s < t_in:
                o[:, s] += dc[:, t, k]


def grad_clip_coef(grad, grad_scale, max_norm, partials, norm_coef, dev_hyper=False, extra_norm=None):
    if dev_hyper:
        grad_scale, max_norm = float(norm_coef[4]), float(norm_coef[5])
    norm = grad.double().norm().item() * grad_scale
    if extra_norm is not None:
        norm = math.sqrt(norm * norm + float(extra_norm[0]) ** 2)
    coef = grad_scale * (min(1.0, max_norm / (norm + 1e-6)) if max_norm > 0 else 1.0)
    norm_coef[0], norm_coef[1] = norm, coef


def adam(param, grad, exp_avg, exp_avg_sq, *, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, step=1,
         norm_coef=None, param_op=None):
    g = grad * (norm_coef[1] if norm_coef is not None else 1.0)
    exp_avg.mul_(betas[0]).add_(g, alpha=1 - betas[0])
    exp_avg_sq.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    wd_lr = weight_decay * lr
    if step == 0:
        step_size, wd_lr = float(norm_coef[2]), float(norm_coef[3])
    else:
        step_size = lr * math.sqrt(1 - betas[1] ** step) / (1 - betas[0] ** step)
    if wd_lr:
        param.mul_(1 - wd_lr)
    param.addcdiv_(exp_avg, exp_avg_sq.sqrt() + eps, value=-step_size)
    if param_op is not None:
        param_op.copy_(param.to(param_op.dtype))


class _FakeLib:
    @staticmethod
    def load():
        class L:
            @staticmethod
            def mm_sumsq_blocks():
                return 4
        return L


_lib = _FakeLib


def _keep_scale(numel, p, seed, site, seed_dev=None):
    """The emulated counter-based mask of a site over `numel` elements (row-major), scaled by 1 / (1 - p)."""
    if seed_dev is not None:
        seed = int(seed) + int(seed_dev.view(-1)[0])
    g = torch.Generator().manual_seed((int(seed) * 1000003 + int(site)) % (2 ** 63 - 1))
    return (torch.rand(numel, generator=g) >= p).float() / (1.0 - p)


def dropout(x, out, p, seed, site, resid=None, seed_dev=None):
    if seed_dev is not None:
        seed = int(seed) + int(seed_dev.view(-1)[0])
    g = torch.Generator().manual_seed((int(seed) * 1000003 + int(site)) % (2 ** 63 - 1))
    keep = torch.rand(x.numel(), generator=g) >= p
    v = (x.reshape(-1).float() * keep / (1.0 - p))
    if resid is not None:
        v = v + resid.reshape(-1)
    out.view(-1).copy_(v.to(out.dtype))
