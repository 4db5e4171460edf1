"""Thin Python wrappers over the C-ABI kernels: shape/dtype checks, pointer passing, current stream.

PyTorch is used only for device memory and the stream handle; nothing here computes with torch ops.
Every function raises if the CUDA library is missing (``_lib.load``) or the tensors are not on a GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

EPI_OP, EPI_RELU_OP, EPI_RESID_F32, EPI_GLU_OP, EPI_GLU_POS_F32, EPI_F32_OP, EPI_GATE, EPI_F32, EPI_MASK_OP = range(9)
EPI_NAMES = ["op", "relu_op", "resid_f32", "glu_op", "glu_pos_f32", "f32_op", "gate", "f32", "mask_op"]
_DT = {torch.bfloat16: 0, torch.float16: 1}

# Launch counter: bench.py reports how many of OUR kernels ran inside the timed region.
launch_count = 0


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DT[dt]
    except KeyError:
        raise TypeError(f"operand dtype must be bfloat16 or float16, got {dt}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("kernel operands must live on a CUDA device (no CPU fallback exists)")
    return t.data_ptr()


# Optional per-launch device timing (bench.py's instrumented pass): when `timing` is a list, every wrapper
# brackets its launch with CUDA events on the launching stream and appends (name, start, end, work) to it.
timing = None
scope = ""      # set by the engines around a phase: timed launches are reported as "name@scope"


class _Launch:
    __slots__ = ("name", "work", "e0")

    def __init__(self, name: str, work: float = 0.0):
        self.name, self.work, self.e0 = name, work, None

    def __enter__(self):
        if timing is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        global launch_count
        launch_count += 1
        if timing is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            timing.append((self.name + ("@" + scope if scope else ""), self.e0, e1, self.work))
        return False


def fbank_tables(device) -> torch.Tensor:
    lib = _lib.load()
    n = lib.mm_fbank_table_floats()
    host = torch.zeros(n, dtype=torch.float32)
    _lib.check(lib.mm_fbank_build_tables(host.data_ptr()), "mm_fbank_build_tables")
    return host.to(device)


def fbank(wav: torch.Tensor, n_samples: torch.Tensor, feats: torch.Tensor, tables: torch.Tensor) -> None:
    """wav [B, N] fp32 (x 2**15) or int16 PCM, n_samples [B] int64 -> feats [B, m, 80] raw log-mel (valid frames)."""
    assert wav.dtype in (torch.float32, torch.int16) and wav.dim() == 2 and wav.stride(1) == 1
    assert n_samples.dtype == torch.int64 and feats.dtype == torch.float32 and feats.is_contiguous()
    B, m = feats.shape[0], feats.shape[1]
    assert feats.shape[2] == 80
    lib = _lib.load()
    fn = lib.mm_fbank_f32 if wav.dtype == torch.float32 else lib.mm_fbank_i16
    with _Launch("fbank", float(wav.element_size() * wav.numel() + 4 * feats.numel())):   # algorithmic bytes
        _lib.check(fn(_ptr(wav), _ptr(n_samples), B, wav.stride(0), _ptr(feats), m, _ptr(tables), _stream()),
                   "mm_fbank")


def cmvn_stats(feats: torch.Tensor, lens: torch.Tensor, lengths_are_samples: bool, mean_std: torch.Tensor) -> None:
    """Per-utterance (mean, std) [B, 2, 80] with the reference's sequential fp32 arithmetic."""
    assert feats.dtype == torch.float32 and feats.is_contiguous() and feats.shape[2] == 80
    assert mean_std.dtype == torch.float32 and mean_std.is_contiguous() and mean_std.numel() == feats.shape[0] * 160
    lib = _lib.load()
    with _Launch("cmvn_stats", float(4 * feats.numel())):
        _lib.check(lib.mm_cmvn_stats(_ptr(feats), _ptr(lens), int(lengths_are_samples), feats.shape[0],
                                     feats.shape[1], _ptr(mean_std), _stream()), "mm_cmvn_stats")


def cmvn_apply(feats: torch.Tensor, stats: Optional[torch.Tensor], lens: torch.Tensor, lengths_are_samples: bool,
               out_f32: Optional[torch.Tensor], out_op: Optional[torch.Tensor], op_row_offset: int = 0,
               spec_masks: Optional[torch.Tensor] = None, n_fmask: int = 0, n_tmask: int = 0,
               mask_value: float = 0.0) -> None:
    """CMVN + zero padding (+ SpecAugment: spec_masks [B, 2 (n_fmask + n_tmask)] int32, see the header)."""
    if spec_masks is not None:
        assert spec_masks.dtype == torch.int32 and spec_masks.is_contiguous()
        assert spec_masks.numel() == feats.shape[0] * 2 * (n_fmask + n_tmask)
    assert feats.dtype == torch.float32 and feats.is_contiguous() and feats.shape[2] == 80
    assert stats is None or stats.dtype == torch.float32
    B, m = feats.shape[0], feats.shape[1]
    op_frames, dt = 0, 0
    if out_op is not None:
        assert out_op.is_contiguous() and out_op.shape[0] == B and out_op.shape[2] == 80
        op_frames, dt = out_op.shape[1], dtype_code(out_op.dtype)
    if out_f32 is not None:
        assert out_f32.is_contiguous() and out_f32.shape == feats.shape
    lib = _lib.load()
    work = 4.0 * feats.numel() + (4.0 * out_f32.numel() if out_f32 is not None else 0.0) + \
        (2.0 * out_op.numel() if out_op is not None else 0.0)
    with _Launch("cmvn_apply", work):
        _lib.check(lib.mm_cmvn_apply_specaug(_ptr(feats), _ptr(stats), _ptr(lens), int(lengths_are_samples), B, m,
                                             _ptr(out_f32), _ptr(out_op), op_frames, op_row_offset, dt, _ptr(spec_masks),
                                             n_fmask, n_tmask, mask_value, _stream()), "mm_cmvn_apply")


def seq_lens(lens: torch.Tensor, lengths_are_samples: bool, n_layers: int, out: torch.Tensor) -> None:
    assert lens.dtype == torch.int64 and out.dtype == torch.int32
    lib = _lib.load()
    with _Launch("seq_lens"):
        _lib.check(lib.mm_seq_lens(_ptr(lens), int(lengths_are_samples), lens.numel(), n_layers, _ptr(out),
                                   _stream()), "mm_seq_lens")


def seq_lens_mask(lens: torch.Tensor, lengths_are_samples: bool, n_layers: int, out: torch.Tensor, mask: torch.Tensor) -> None:
    """out [B] int32 = subsampled lengths, mask [B, T] bool = t >= out[b]: ``seq_lens`` + ``padding_mask`` in one launch."""
    assert lens.dtype == torch.int64 and out.dtype == torch.int32 and mask.dtype == torch.bool and mask.is_contiguous()
    assert mask.shape[0] == lens.numel() == out.numel()
    lib = _lib.load()
    with _Launch("seq_lens_mask"):
        _lib.check(lib.mm_seq_lens_mask(_ptr(lens), int(lengths_are_samples), lens.numel(), n_layers, mask.shape[1],
                                        _ptr(out), _ptr(mask), _stream()), "mm_seq_lens_mask")


def padding_mask(seq_lens: torch.Tensor, T: int, out: torch.Tensor) -> None:
    """out [B, T] bool = t >= seq_lens[b]."""
    assert seq_lens.dtype == torch.int32 and out.dtype == torch.bool and out.is_contiguous()
    assert out.shape == (seq_lens.numel(), T)
    lib = _lib.load()
    with _Launch("padding_mask"):
        _lib.check(lib.mm_padding_mask(_ptr(seq_lens), seq_lens.numel(), T, _ptr(out), _stream()), "mm_padding_mask")


def gemm(*, a0: torch.Tensor, w: torch.Tensor, rows: int, n: int, k: int, mode: int, out0: torch.Tensor,
         a0_ld: int, out0_ld: int, batches: int = 1, a0_bs: int = 0, a1: Optional[torch.Tensor] = None,
         a1_ld: int = 0, a1_bs: int = 0, k_split: int = 0, w_ld: Optional[int] = None, w_bs: int = 0,
         w_batched: bool = False, bias: Optional[torch.Tensor] = None, scale: float = 1.0, scale_cols: int = 0,
         out0_bs: int = 0, out1: Optional[torch.Tensor] = None, out1_ld: int = 0, out1_bs: int = 0,
         aux0: Optional[torch.Tensor] = None, aux1: Optional[torch.Tensor] = None, aux_ld: int = 0,
         rows_per_seq: int = 0, out_tbc: bool = False, n_seqs: int = 0, out_row_offset: int = 0,
         vt: Optional[torch.Tensor] = None, vt_col0: int = 0, vt_rows: int = 0, vt_ld: int = 0,
         pos: Optional[torch.Tensor] = None, seq_lens: Optional[torch.Tensor] = None, block_n: int = 0,
         a_mn: bool = False, w_mn: bool = False, a_kbatch: bool = False, w_kbatch: bool = False, a_hm: bool = False,
         w_hm: bool = False, out_hm: bool = False, heads: int = 0, head_stride: int = 0, a_k_total: int = 0,
         w_k_total: int = 0, drop=None) -> None:
    """acc = A @ W^T with a fused epilogue; see ``mm_gemm_args`` in include/mms2ut_b200.h."""
    if a0.dtype != w.dtype or (a1 is not None and a1.dtype != w.dtype):
        raise TypeError("A and W must share the 16-bit operand dtype")
    for t in (bias, aux0 if mode != EPI_MASK_OP else None, aux1, pos):
        if t is not None and t.dtype != torch.float32:
            raise TypeError("bias/aux/pos must be float32")
    if mode == EPI_MASK_OP and (aux0 is None or aux0.dtype != w.dtype):
        raise TypeError("MASK_OP: aux0 is the kept 16-bit activation")
    if seq_lens is not None and seq_lens.dtype != torch.int32:
        raise TypeError("seq_lens must be int32")
    g = _lib.GemmArgs()
    g.a0, g.a1, g.w = _ptr(a0), _ptr(a1), _ptr(w)
    g.a0_ld, g.a0_bs, g.a1_ld, g.a1_bs = a0_ld, a0_bs, a1_ld, a1_bs
    g.w_ld, g.w_bs = (k if w_ld is None else w_ld), w_bs
    g.rows, g.batches, g.n, g.k, g.k_split, g.w_batched = rows, batches, n, k, k_split, int(w_batched)
    g.dtype, g.mode, g.block_n = dtype_code(w.dtype), mode, block_n
    g.bias, g.scale, g.scale_cols = _ptr(bias), scale, scale_cols
    g.out0, g.out0_ld, g.out0_bs = _ptr(out0), out0_ld, out0_bs
    g.out1, g.out1_ld, g.out1_bs = _ptr(out1), out1_ld, out1_bs
    g.aux0, g.aux1, g.aux_ld = _ptr(aux0), _ptr(aux1), aux_ld
    g.rows_per_seq, g.out_tbc, g.n_seqs, g.out_row_offset = rows_per_seq, int(out_tbc), n_seqs, out_row_offset
    g.vt, g.vt_col0, g.vt_rows, g.vt_ld = _ptr(vt), vt_col0, vt_rows, vt_ld
    g.pos, g.seq_lens = _ptr(pos), _ptr(seq_lens)
    g.a_mn, g.w_mn, g.a_kbatch, g.w_kbatch = int(a_mn), int(w_mn), int(a_kbatch), int(w_kbatch)
    g.a_hm, g.w_hm, g.out_hm, g.heads, g.head_stride = int(a_hm), int(w_hm), int(out_hm), heads, head_stride
    g.a_k_total, g.w_k_total = a_k_total, w_k_total
    if drop is not None and drop[0] > 0:      # RELU_OP: activation dropout in the epilogue, drop = (p, seed, seed_dev, site)
        g.drop_p, g.drop_seed, g.drop_seed_dev, g.drop_site = drop[0], drop[1] & 0xFFFFFFFFFFFFFFFF, _ptr(drop[2]), drop[3]
    lib = _lib.load()
    with _Launch(f"gemm[{EPI_NAMES[mode]}]", 2.0 * rows * batches * n * k):          # algorithmic FLOPs
        _lib.check(lib.mm_gemm(C.byref(g), _stream()), "mm_gemm")


def gemm_resid_ln(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor,
                  beta: torch.Tensor, h_op: torch.Tensor, h_f32: Optional[torch.Tensor] = None,
                  eps: float = 1e-5, x_out: Optional[torch.Tensor] = None, drop=None) -> None:
    """x_out (default: x, in place) = x + dropout(a @ w.T + bias) (fp32); h_op (and h_f32) = LayerNorm(x_out) * gamma +
    beta.  n = 512 only.  drop = (p, seed, seed_dev, site) or None: the counter-based mask of ``dropout`` over the
    [rows, n] sub-layer output."""
    if x_out is None:
        x_out = x
    assert x_out.dtype == torch.float32 and x_out.is_contiguous() and x_out.shape == x.shape
    rows, k = a.shape
    n = w.shape[0]
    assert a.dtype == w.dtype == h_op.dtype and a.is_contiguous() and w.is_contiguous() and w.shape[1] == k
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape == (rows, n) and h_op.shape == (rows, n)
    assert all(t.dtype == torch.float32 and t.numel() == n for t in (bias, gamma, beta))
    lib = _lib.load()
    p_, seed, seed_dev, site = drop if drop is not None else (0.0, 0, None, 0)
    with _Launch("gemm_resid_ln", 2.0 * rows * n * k):
        _lib.check(lib.mm_gemm_resid_ln_drop(_ptr(a), k, _ptr(w), k, rows, k, n, _ptr(bias), _ptr(x), _ptr(x_out),
                                             _ptr(gamma), _ptr(beta), eps, _ptr(h_op), _ptr(h_f32), float(p_),
                                             seed & 0xFFFFFFFFFFFFFFFF, _ptr(seed_dev), site, dtype_code(w.dtype),
                                             _stream()), "mm_gemm_resid_ln")


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out_op: Optional[torch.Tensor] = None,
              out_f32: Optional[torch.Tensor] = None, eps: float = 1e-5) -> None:
    assert x.dtype == torch.float32 and x.is_contiguous()
    dim = x.shape[-1]
    rows = x.numel() // dim
    assert gamma.dtype == torch.float32 and gamma.numel() == dim and beta.numel() == dim
    dt = dtype_code(out_op.dtype) if out_op is not None else 0
    lib = _lib.load()
    work = 4.0 * x.numel() + (2.0 * x.numel() if out_op is not None else 0.0) + \
        (4.0 * x.numel() if out_f32 is not None else 0.0)
    with _Launch("layernorm", work):
        _lib.check(lib.mm_layernorm(_ptr(x), _ptr(gamma), _ptr(beta), rows, dim, _ptr(out_op), _ptr(out_f32), dt,
                                    eps, _stream()), "mm_layernorm")


def layernorm_gather(store: torch.Tensor, index: Optional[torch.Tensor], rows_per_index: int, gamma: torch.Tensor,
                     beta: torch.Tensor, out_op: torch.Tensor, eps: float = 1e-5) -> None:
    """out_op[r] = LayerNorm(store[index[r // rows_per_index], r % rows_per_index]): store [N, rows_per_index, dim]
    16-bit on the device, index [B] int64 (None: rows in order)."""
    assert store.dtype in _DT and store.is_contiguous() and out_op.is_contiguous()
    dim = store.shape[-1]
    rows = out_op.shape[0]
    assert out_op.shape[1] == dim and gamma.dtype == beta.dtype == torch.float32
    if index is not None:
        assert index.dtype == torch.int64 and index.numel() * rows_per_index == rows
    lib = _lib.load()
    with _Launch("layernorm_gather", float(store.element_size() * rows * dim + out_op.element_size() * rows * dim)):
        _lib.check(lib.mm_layernorm_gather(_ptr(store), dtype_code(store.dtype), _ptr(index), rows_per_index,
                                           _ptr(gamma), _ptr(beta), rows, dim, _ptr(out_op), dtype_code(out_op.dtype),
                                           eps, _stream()), "mm_layernorm_gather")


def self_attention_drop_supported(seq: int) -> bool:
    """Attention dropout inside the forward kernel: every length (single-chunk and chunked kernels alike)."""
    return seq > 0


def self_attention(qkv: torch.Tensor, seq_lens_: torch.Tensor, batch: int, seq: int, heads: int,
                   out: torch.Tensor, lse: Optional[torch.Tensor] = None, drop=None) -> None:
    """qkv [B*T, 3d] (q pre-scaled | k | v) -> out [B*T, d]; head_dim 64.  lse (optional): [B, heads, T] fp32
    log-sum-exp of every query row's scores (kept by the training forward for ``attention_bwd_scores``).
    drop = (p, seed, seed_dev, site): attention dropout on the probabilities inside the kernel (training forward)."""
    assert qkv.dtype == out.dtype and seq_lens_.dtype == torch.int32 and qkv.stride(-1) == 1
    assert lse is None or (lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == batch * heads * seq)
    lib = _lib.load()
    if drop is not None and drop[0] > 0:
        with _Launch("self_attention", 4.0 * batch * seq * seq * heads * 64):
            _lib.check(lib.mm_self_attention_drop(_ptr(qkv), qkv.stride(-2), _ptr(seq_lens_), batch, seq, heads, _ptr(out),
                                                  out.stride(-2), _ptr(lse), float(drop[0]),
                                                  drop[1] & 0xFFFFFFFFFFFFFFFF, _ptr(drop[2]), drop[3],
                                                  dtype_code(qkv.dtype), _stream()), "mm_self_attention_drop")
        return
    with _Launch("self_attention", 4.0 * batch * seq * seq * heads * 64):
        _lib.check(lib.mm_self_attention_lse(_ptr(qkv), qkv.stride(-2), _ptr(seq_lens_), batch, seq, heads, _ptr(out),
                                             out.stride(-2), _ptr(lse), dtype_code(qkv.dtype), _stream()),
                   "mm_self_attention")


def label_smoothed_nll(logits: torch.Tensor, vocab: int, target: torch.Tensor, padding_idx: int, epsilon: float):
    """fairseq ``label_smoothed_nll_loss(log_softmax(logits), target, epsilon, ignore_index=padding_idx)`` summed over
    rows: returns (loss, nll_loss) as 0-d device tensors.  logits fp32 [rows, ld >= vocab], target int64 [rows]."""
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
    assert target.dtype == torch.int64 and target.numel() == logits.shape[0] and target.is_contiguous()
    rows = logits.shape[0]
    row_nll = torch.empty(rows, dtype=torch.float32, device=logits.device)
    row_smooth = torch.empty(rows, dtype=torch.float32, device=logits.device)
    sums = torch.empty(2, dtype=torch.float32, device=logits.device)
    lib = _lib.load()
    with _Launch("label_smoothed_nll", 4.0 * rows * vocab):
        _lib.check(lib.mm_label_smoothed_nll(_ptr(logits), logits.stride(0), vocab, _ptr(target), padding_idx, rows,
                                             _ptr(row_nll), _ptr(row_smooth), _ptr(sums), _stream()),
                   "mm_label_smoothed_nll")
    eps_i = epsilon / (vocab - 1)
    return (1.0 - epsilon - eps_i) * sums[0] + eps_i * sums[1], sums[0]


def embed_tokens(tokens: torch.Tensor, padding_idx: int, table: torch.Tensor, scale: float, pos_table: torch.Tensor,
                 out: torch.Tensor) -> None:
    """out [B*L, d] fp32 = scale * table[tokens] + pos_table[fairseq positions(tokens)]."""
    assert tokens.dtype == torch.int64 and tokens.dim() == 2 and tokens.is_contiguous()
    assert table.dtype == pos_table.dtype == out.dtype == torch.float32
    assert table.is_contiguous() and pos_table.is_contiguous() and out.is_contiguous()
    B, L = tokens.shape
    lib = _lib.load()
    with _Launch("embed_tokens", 8.0 * B * L * table.shape[1]):
        _lib.check(lib.mm_embed_tokens(_ptr(tokens), padding_idx, _ptr(table), table.shape[0], scale, _ptr(pos_table),
                                       pos_table.shape[0], B, L, table.shape[1], _ptr(out), _stream()),
                   "mm_embed_tokens")


def attention(q: torch.Tensor, q_col0: int, q_len: int, k: torch.Tensor, k_col0: int, v: torch.Tensor, v_col0: int,
              kv_len: int, kv_lens: Optional[torch.Tensor], batch: int, heads: int, out: torch.Tensor,
              causal: bool = False, lse: Optional[torch.Tensor] = None, drop=None) -> None:
    """General attention core (see ``mm_attention``): q [batch*q_len, ld] (pre-scaled), k / v [batch*kv_len, ld],
    head h in columns [col0 + 64 h, col0 + 64 h + 64); out [batch*q_len, heads*64].  drop = (p, seed, seed_dev, site):
    attention dropout inside the kernel (``mm_attention_drop``; mask index ((b H + h) Lp + q) Tp + k)."""
    assert q.dtype == k.dtype == v.dtype == out.dtype and q.dtype in _DT
    assert all(t.dim() == 2 and t.stride(1) == 1 for t in (q, k, v, out))
    assert q.shape[0] == batch * q_len and k.shape[0] == batch * kv_len and v.shape[0] == batch * kv_len
    if kv_lens is not None:
        assert kv_lens.dtype == torch.int32 and kv_lens.numel() == batch
    lib = _lib.load()
    with _Launch("attention", 4.0 * batch * heads * q_len * kv_len * 64 * (0.5 if causal else 1.0)):
        assert lse is None or (lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == batch * heads * q_len)
        p_, seed, seed_dev, site = drop if (drop is not None and drop[0] > 0) else (0.0, 0, None, 0)
        _lib.check(lib.mm_attention_drop(_ptr(q), q.stride(0), q_col0, q_len, _ptr(k), k.stride(0), k_col0, _ptr(v),
                                         v.stride(0), v_col0, kv_len, _ptr(kv_lens), batch, heads, int(causal), _ptr(out),
                                         out.stride(0), _ptr(lse), float(p_), seed & 0xFFFFFFFFFFFFFFFF, _ptr(seed_dev),
                                         site, dtype_code(q.dtype), _stream()), "mm_attention")


def attention_bwd_scores(q: torch.Tensor, q_col0: int, q_len: int, k: torch.Tensor, k_col0: int, v: torch.Tensor,
                         v_col0: int, kv_len: int, kv_lens: Optional[torch.Tensor], batch: int, heads: int, dout: torch.Tensor,
                         out: torch.Tensor, lse: torch.Tensor, probs: torch.Tensor, dscores: torch.Tensor,
                         causal: bool = False) -> None:
    """P = exp(q k^T - lse) (masked), dS = P o (dO v^T - rowsum(dO o O)) -> probs / dscores [batch*heads, Lp, Tp] 16-bit
    (see ``mm_attention_bwd_scores``); q / k / v as in ``attention``; dout / out [batch*q_len, heads*64]."""
    assert q.dtype == k.dtype == v.dtype == dout.dtype == out.dtype == probs.dtype == dscores.dtype and q.dtype in _DT
    assert all(t.dim() == 2 and t.stride(1) == 1 for t in (q, k, v, dout, out))
    assert probs.dim() == 3 and probs.is_contiguous() and dscores.shape == probs.shape and dscores.is_contiguous()
    assert probs.shape[0] == batch * heads and probs.shape[1] >= q_len and probs.shape[2] >= kv_len
    assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == batch * heads * q_len
    if kv_lens is not None:
        assert kv_lens.dtype == torch.int32 and kv_lens.numel() == batch
    lib = _lib.load()
    with _Launch("attention_bwd_scores", 4.0 * batch * heads * q_len * kv_len * 64 * (0.5 if causal else 1.0)):
        _lib.check(lib.mm_attention_bwd_scores(_ptr(q), q.stride(0), q_col0, q_len, _ptr(k), k.stride(0), k_col0, _ptr(v),
                                               v.stride(0), v_col0, kv_len, _ptr(kv_lens), batch, heads, int(causal),
                                               _ptr(dout), dout.stride(0), _ptr(out), out.stride(0), _ptr(lse),
                                               _ptr(probs), _ptr(dscores), probs.stride(1), probs.stride(0),
                                               dtype_code(q.dtype), _stream()), "mm_attention_bwd_scores")


def cross_attention(q: torch.Tensor, q_len: int, k: torch.Tensor, k_col0: int, v: torch.Tensor, v_col0: int, kv_len: int,
                    batch: int, d_model: int, out: torch.Tensor, key_mask: Optional[torch.Tensor] = None,
                    lse: Optional[torch.Tensor] = None, kv_batch_stride: int = 0) -> None:
    """Fused speech -> image attention for one head of width d_model (see ``mm_cross_attention``): q [batch*q_len, ld]
    (pre-scaled), k / v rows [batch][kv_len][ld] with their d_model-wide blocks at *_col0, key_mask [batch, >= kv_len]
    uint8 (non-zero = masked out); out [batch*q_len, d_model]; lse [batch*q_len] fp32 (optional)."""
    assert q.dtype == k.dtype == v.dtype == out.dtype and q.dtype in _DT
    assert all(t.dim() == 2 and t.stride(1) == 1 for t in (q, k, v, out))
    assert q.shape[0] == batch * q_len and out.shape[0] == batch * q_len
    if key_mask is not None:
        assert key_mask.dtype == torch.uint8 and key_mask.dim() == 2 and key_mask.stride(1) == 1
        assert key_mask.shape[0] == batch and key_mask.shape[1] >= kv_len
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == batch * q_len
    lib = _lib.load()
    with _Launch("cross_attention", 4.0 * batch * q_len * kv_len * d_model):
        _lib.check(lib.mm_cross_attention(_ptr(q), q.stride(0), q_len, _ptr(k), k.stride(0), k_col0, _ptr(v), v.stride(0),
                                          v_col0, kv_len, kv_batch_stride, _ptr(key_mask),
                                          key_mask.stride(0) if key_mask is not None else 0, batch, d_model, _ptr(out),
                                          out.stride(0), _ptr(lse), dtype_code(q.dtype), _stream()), "mm_cross_attention")


def softmax_rows(scores: torch.Tensor, ld_in: int, rows: int, n_keys: int, probs: torch.Tensor, ld_out: int,
                 key_mask: Optional[torch.Tensor] = None, rows_per_seq: int = 0) -> None:
    assert scores.dtype == torch.float32
    if key_mask is not None:
        assert key_mask.dtype == torch.uint8 and key_mask.is_contiguous()
    lib = _lib.load()
    with _Launch("softmax_rows", 4.0 * rows * n_keys + 2.0 * rows * ld_out):
        _lib.check(lib.mm_softmax_rows(_ptr(scores), ld_in, rows, n_keys, _ptr(key_mask), rows_per_seq, _ptr(probs),
                                       ld_out, dtype_code(probs.dtype), _stream()), "mm_softmax_rows")


def mask_scores(scores: torch.Tensor, ld: int, rows: int, n_keys: int, key_mask: torch.Tensor, rows_per_seq: int) -> None:
    """scores[r, k] = -inf where key_mask[r // rows_per_seq, k] != 0 (the reference's masked_fill, fuse.py:88-91), in place."""
    assert scores.dtype == torch.float32 and key_mask.dtype == torch.uint8 and key_mask.dim() == 2
    assert key_mask.stride(1) == 1 and key_mask.shape[1] >= n_keys and key_mask.shape[0] * rows_per_seq == rows
    lib = _lib.load()
    with _Launch("mask_scores", 1.0 * rows * n_keys):
        _lib.check(lib.mm_mask_scores(_ptr(scores), ld, rows, n_keys, _ptr(key_mask), key_mask.stride(0), rows_per_seq,
                                      _stream()), "mm_mask_scores")


def convert(x: torch.Tensor, out: torch.Tensor) -> None:
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous() and out.numel() == x.numel()
    lib = _lib.load()
    with _Launch("convert", 6.0 * x.numel()):
        _lib.check(lib.mm_convert_f32(_ptr(x), _ptr(out), x.numel(), dtype_code(out.dtype), _stream()),
                   "mm_convert_f32")


# ----------------------------------------------------------------------------------------------------------
# training-step variant (csrc/backward.cu): see include/mms2ut_b200.h
# ----------------------------------------------------------------------------------------------------------
def pack_t(x: torch.Tensor, *, rows: int, cols: int, in_ld: int, out_n: Optional[torch.Tensor] = None, n_ld: int = 0,
           out_t: Optional[torch.Tensor] = None, t_ld: int = 0, t_cols_pad: int = 0, batches: int = 1, nb1: int = 1,
           in_bs0: int = 0, in_bs1: int = 0, n_bs0: int = 0, n_bs1: int = 0, t_bs0: int = 0, t_bs1: int = 0,
           mask: Optional[torch.Tensor] = None, mask_ld: int = 0, scale: float = 1.0) -> None:
    """16-bit straight (out_n) and/or transposed (out_t, zero-padded to t_cols_pad columns) copies of x (fp32/16-bit)."""
    outs = [t for t in (out_n, out_t) if t is not None]
    assert outs and all(t.dtype == outs[0].dtype for t in outs)
    assert x.dtype == torch.float32 or x.dtype == outs[0].dtype
    if mask is not None:
        assert mask.dtype == outs[0].dtype
    lib = _lib.load()
    with _Launch("pack_t", float(x.element_size() * rows * cols * batches +
                                 2 * rows * cols * batches * len(outs))):
        _lib.check(lib.mm_pack_t(_ptr(x), int(x.dtype == torch.float32), in_ld, in_bs0, in_bs1, nb1, _ptr(mask), mask_ld,
                                 rows, cols, batches, scale, _ptr(out_n), n_ld, n_bs0, n_bs1, _ptr(out_t), t_ld, t_bs0,
                                 t_bs1, t_cols_pad, dtype_code(outs[0].dtype), _stream()), "mm_pack_t")


def rowsum(x: torch.Tensor, ld: int, rows: int, cols: int, out: torch.Tensor, accumulate: bool = False) -> None:
    assert out.dtype == torch.float32 and out.numel() >= rows
    lib = _lib.load()
    with _Launch("rowsum", 2.0 * rows * cols):
        _lib.check(lib.mm_rowsum(_ptr(x), ld, rows, cols, _ptr(out), int(accumulate), dtype_code(x.dtype), _stream()),
                   "mm_rowsum")


def reduce_partials(part: torch.Tensor, n_partials: int, stride: int, n: int, out: torch.Tensor,
                    accumulate: bool = False, part_offset: int = 0) -> None:
    assert part.dtype == torch.float32 and out.dtype == torch.float32 and out.numel() >= n
    lib = _lib.load()
    with _Launch("reduce_partials", 4.0 * n * n_partials):
        _lib.check(lib.mm_reduce_partials(_ptr(part) + 4 * part_offset, n_partials, stride, n, _ptr(out),
                                          int(accumulate), _stream()), "mm_reduce_partials")


def reduce_partials_many(jobs) -> None:
    """jobs: list of (part, n_partials, stride, n, out, accumulate): all reductions in one launch per 16 jobs."""
    if not jobs:
        return
    arr = (_lib.ReduceJob * len(jobs))()
    work = 0.0
    for a, (part, S, stride, n, out, acc) in zip(arr, jobs):
        assert part.dtype == torch.float32 and out.dtype == torch.float32 and out.numel() >= n
        a.part, a.out, a.stride, a.n, a.n_partials, a.accumulate = _ptr(part), _ptr(out), stride, n, S, int(acc)
        work += 4.0 * n * S
    lib = _lib.load()
    with _Launch("reduce_partials_many", work):
        _lib.check(lib.mm_reduce_partials_many(arr, len(jobs), _stream()), "mm_reduce_partials_many")


def attention_bwd_fused(qkv: torch.Tensor, seq_len: int, kv_lens: Optional[torch.Tensor], batch: int, heads: int,
                        dout: torch.Tensor, out: torch.Tensor, lse: torch.Tensor, dqkv: torch.Tensor, drop=None) -> None:
    """dq | dk | dv [batch*seq_len, 3 * heads*64] from q | k | v (same layout), dout, out and lse in one kernel; seq_len
    <= 256 (``mm_attention_bwd_fused``).  drop = (p, seed, seed_dev, site): the attention-dropout mask the forward
    kernel applied (``self_attention(..., drop=)``) is regenerated."""
    d = heads * 64
    assert qkv.dtype == dout.dtype == out.dtype == dqkv.dtype and qkv.dtype in _DT
    assert all(t.dim() == 2 and t.stride(1) == 1 for t in (qkv, dout, out, dqkv))
    assert qkv.shape[1] >= 3 * d and dqkv.shape[1] >= 3 * d and seq_len <= 256
    assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == batch * heads * seq_len
    if kv_lens is not None:
        assert kv_lens.dtype == torch.int32 and kv_lens.numel() == batch
    lib = _lib.load()
    p_, seed, seed_dev, site = drop if (drop is not None and drop[0] > 0) else (0.0, 0, None, 0)
    with _Launch("attention_bwd_fused", 10.0 * batch * heads * seq_len * seq_len * 64):
        _lib.check(lib.mm_attention_bwd_fused_drop(_ptr(qkv), qkv.stride(0), 0, d, 2 * d, seq_len, _ptr(kv_lens), batch,
                                                   heads, _ptr(dout), dout.stride(0), _ptr(out), out.stride(0), _ptr(lse),
                                                   _ptr(dqkv), dqkv.stride(0), float(p_), seed & 0xFFFFFFFFFFFFFFFF,
                                                   _ptr(seed_dev), site, dtype_code(qkv.dtype), _stream()),
                   "mm_attention_bwd_fused")


def attention_bwd_general_scratch_floats(kv_len: int) -> int:
    return int(_lib.load().mm_attention_bwd_general_scratch_floats(kv_len))


def attention_bwd_general(q: torch.Tensor, q_len: int, k: torch.Tensor, v: torch.Tensor, kv_len: int,
                          kv_lens: Optional[torch.Tensor], batch: int, heads: int, dout: torch.Tensor, out: torch.Tensor,
                          lse: torch.Tensor, dq: torch.Tensor, dk: torch.Tensor, dv: torch.Tensor, scratch: torch.Tensor,
                          causal: bool = False, drop=None) -> None:
    """dq / dk / dv from q / k / v (2-D token-major views whose first column is head 0's), dout, out and lse in one kernel
    for any lengths, causal or not (``mm_attention_bwd_general``); scratch: fp32, attention_bwd_general_scratch_floats.
    drop = (p, seed, seed_dev, site): the attention-dropout mask of the forward kernel is regenerated."""
    ts = (q, k, v, dout, out, dq, dk, dv)
    assert all(t.dtype == q.dtype for t in ts) and q.dtype in _DT
    assert all(t.dim() == 2 and t.stride(1) == 1 for t in ts)
    assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == batch * heads * q_len
    assert scratch.dtype == torch.float32 and scratch.numel() >= attention_bwd_general_scratch_floats(kv_len)
    if kv_lens is not None:
        assert kv_lens.dtype == torch.int32 and kv_lens.numel() == batch
    lib = _lib.load()
    with _Launch("attention_bwd_general", 10.0 * batch * heads * q_len * kv_len * 64 * (0.5 if causal else 1.0)):
        p_, seed, seed_dev, site = drop if (drop is not None and drop[0] > 0) else (0.0, 0, None, 0)
        _lib.check(lib.mm_attention_bwd_general_drop(_ptr(q), q.stride(0), 0, q_len, _ptr(k), k.stride(0), 0, _ptr(v),
                                                     v.stride(0), 0, kv_len, _ptr(kv_lens), batch, heads, int(causal),
                                                     _ptr(dout), dout.stride(0), _ptr(out), out.stride(0), _ptr(lse),
                                                     _ptr(dq), dq.stride(0), 0, _ptr(dk), dk.stride(0), 0, _ptr(dv),
                                                     dv.stride(0), 0, _ptr(scratch), float(p_),
                                                     seed & 0xFFFFFFFFFFFFFFFF, _ptr(seed_dev), site,
                                                     dtype_code(q.dtype), _stream()), "mm_attention_bwd_general")


def heads_gemm(a: torch.Tensor, a_ld: int, a_bs: int, transposed: bool, w: torch.Tensor, w_ld: int, w_bs: int,
               out: torch.Tensor, out_ld: int, out_bs: int, rows: int, k: int, batch: int, heads: int,
               scale: float = 1.0) -> None:
    """Per (sequence, head), head_dim 64: out[b][r][64 h + c] = scale * sum_j A_bh[r, j] w[b][j][64 h + c] with A_bh the
    [rows, k] matrix number b * heads + h of ``a`` (read transposed when ``transposed``); w / out are (views of)
    token-major tensors whose first column is head 0's (``mm_heads_gemm``: dV = P^T dO, dK = dS^T q, dQ = dS k)."""
    if not (a.dtype == w.dtype == out.dtype):
        raise TypeError("heads_gemm: a, w and out share the 16-bit operand dtype")
    lib = _lib.load()
    with _Launch("heads_gemm", 2.0 * batch * heads * rows * k * 64):
        _lib.check(lib.mm_heads_gemm(_ptr(a), a_ld, a_bs, int(transposed), _ptr(w), w_ld, w_bs, 0, _ptr(out), out_ld,
                                     out_bs, 0, rows, k, batch, heads, scale, dtype_code(w.dtype), _stream()),
                   "mm_heads_gemm")


def wgrad_grouped(groups, tokens: int, accumulate: bool = False) -> None:
    """groups: list of (dy [tokens, n_out] 16-bit, dy_ld, x [tokens, k_in] 16-bit, x_ld, out fp32, out_ld, n_out, k_in,
    bias fp32 [n_out] or None): out[n, k] (+)= sum_t dy[t, n] x[t, k] and bias[n] (+)= sum_t dy[t, n] (x = out = None,
    k_in = 0: the bias only) for every group in one launch (``mm_wgrad_grouped``: the groups' output tiles
    share the persistent grid, full token contraction per tile, no partials); out_ld = row stride of out (a column
    block of a wider gradient is out = wide.view(-1)[col:], out_ld = wide row length).  A group may carry a 10th
    element, its own token count (groups of different lengths share the launch; the library balances the tile order)."""
    lib = _lib.load()
    for base in range(0, len(groups), _lib.WGRAD_MAX_GROUPS):
        chunk = groups[base:base + _lib.WGRAD_MAX_GROUPS]
        arr = (_lib.WgradGroup * len(chunk))()
        work = 0.0
        for a, grp in zip(arr, chunk):
            dy, dy_ld, x, x_ld, out, out_ld, n_out, k_in, bias = grp[:9]
            a.tokens = int(grp[9]) if len(grp) > 9 else 0
            tokens_g = a.tokens or tokens
            if k_in > 0:
                if dy.dtype != x.dtype or out.dtype != torch.float32:
                    raise TypeError("wgrad_grouped: dy / x share the 16-bit operand dtype, out is float32")
                assert out.numel() >= (n_out - 1) * out_ld + k_in
            if bias is not None:
                assert bias.dtype == torch.float32 and bias.numel() >= n_out
            a.dy, a.x, a.out, a.bias = _ptr(dy), _ptr(x), _ptr(out), _ptr(bias)
            a.dy_ld, a.x_ld, a.out_ld, a.n_out, a.k_in = dy_ld, x_ld, out_ld, n_out, k_in
            work += 2.0 * tokens_g * n_out * k_in
        with _Launch("wgrad_grouped", work):
            _lib.check(lib.mm_wgrad_grouped(arr, len(chunk), tokens, int(accumulate), dtype_code(chunk[0][0].dtype),
                                            _stream()), "mm_wgrad_grouped")


def layernorm_bwd_blocks() -> int:
    return _lib.load().mm_layernorm_bwd_blocks()


def colsum_blocks(rows: int) -> int:
    return (rows + 127) // 128


def colsum(x: torch.Tensor, ld: int, rows: int, cols: int, partials: torch.Tensor, period: int = 0,
           valid: int = 0) -> int:
    """partials [colsum_blocks(rows), cols] = per-chunk column sums of the 16-bit matrix x [rows, cols]; returns the
    number of partial rows (sum them with ``reduce_partials``)."""
    nb = colsum_blocks(rows)
    assert partials.dtype == torch.float32 and partials.numel() >= nb * cols
    lib = _lib.load()
    with _Launch("colsum", 2.0 * rows * cols):
        _lib.check(lib.mm_colsum(_ptr(x), ld, rows, cols, period, valid, _ptr(partials), dtype_code(x.dtype), _stream()),
                   "mm_colsum")
    return nb


def layernorm_bwd(x: torch.Tensor, gamma: torch.Tensor, dy: torch.Tensor, partials: torch.Tensor,
                  dx: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None, eps: float = 1e-5,
                  dx_op: Optional[torch.Tensor] = None, drop=None) -> None:
    """dx = resid + LayerNorm'(dy) (dx_op: its 16-bit copy); partials [blocks, 2, dim] per-block (dgamma, dbeta) sums.
    drop = (p, seed, seed_dev, site): dx_op = dx o keep / (1 - p), the gradient entering the dropped sub-layer branch
    that consumes it (dx, the residual path, stays unmasked)."""
    assert dx_op is None or (dx_op.is_contiguous() and dx_op.numel() == x.numel())
    dim = x.shape[-1]
    rows = x.numel() // dim
    for t in (x, gamma, dy, partials, dx, resid):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    assert dy.numel() == x.numel() and partials.numel() >= layernorm_bwd_blocks() * 2 * dim
    lib = _lib.load()
    with _Launch("layernorm_bwd", 4.0 * x.numel() * (2 + (dx is not None) + (resid is not None))):
        p_, seed, seed_dev, site = drop if (drop is not None and drop[0] > 0) else (0.0, 0, None, 0)
        _lib.check(lib.mm_layernorm_bwd_drop(_ptr(x), _ptr(gamma), _ptr(dy), rows, dim, eps, _ptr(resid), _ptr(dx),
                                             _ptr(partials), _ptr(dx_op), float(p_), seed & 0xFFFFFFFFFFFFFFFF,
                                             _ptr(seed_dev), site,
                                             dtype_code(dx_op.dtype) if dx_op is not None else 0, _stream()),
                   "mm_layernorm_bwd")


def gemm_ln_bwd_partial_rows(rows: int) -> int:
    return int(_lib.load().mm_gemm_ln_bwd_partial_rows(rows))


def gemm_ln_bwd(dy: torch.Tensor, w: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, g: torch.Tensor,
                g_op: torch.Tensor, partials: torch.Tensor, eps: float = 1e-5, drop=None) -> int:
    """g += LayerNorm'(dy @ w; x, gamma) in place, g_op = its 16-bit copy (x the dropout mask ``drop`` = (p, seed,
    seed_dev, site) of the branch that consumes it), partials [rows_out, 2, 512] = (dgamma, dbeta) partial sums
    (``mm_gemm_ln_bwd``: dgrad of a Linear after a LayerNorm + the LayerNorm backward + the residual add; d_model 512).
    dy [rows, k] 16-bit (any row stride), w [k, 512] = the Linear's weight as stored.  Returns the number of partial
    rows written."""
    rows, k = dy.shape
    assert dy.dtype == w.dtype == g_op.dtype and dy.dtype in _DT and dy.stride(1) == 1
    assert w.shape == (k, 512) and w.stride(1) == 1
    for t in (x, g):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == rows * 512
    assert gamma.dtype == torch.float32 and gamma.numel() == 512 and g_op.is_contiguous() and g_op.numel() == rows * 512
    nrows = gemm_ln_bwd_partial_rows(rows)
    assert partials.dtype == torch.float32 and partials.numel() >= nrows * 1024
    p_, seed, seed_dev, site = drop if (drop is not None and drop[0] > 0) else (0.0, 0, None, 0)
    lib = _lib.load()
    with _Launch("gemm_ln_bwd", 2.0 * rows * 512 * k):
        _lib.check(lib.mm_gemm_ln_bwd(_ptr(dy), dy.stride(0), _ptr(w), w.stride(0), rows, k, _ptr(x), _ptr(gamma), eps,
                                      _ptr(g), _ptr(g_op), _ptr(partials), float(p_), seed & 0xFFFFFFFFFFFFFFFF,
                                      _ptr(seed_dev), site, dtype_code(dy.dtype), _stream()), "mm_gemm_ln_bwd")
    return nrows


def softmax_bwd(scores: torch.Tensor, dprobs: Optional[torch.Tensor], ld_in: int, rows: int, rows_per_batch: int,
                n_keys: int, dscores: Optional[torch.Tensor], ld_out: int, probs: Optional[torch.Tensor] = None,
                kv_lens: Optional[torch.Tensor] = None, heads: int = 1, valid_rows: int = 0,
                causal: bool = False, ld_dprobs: Optional[int] = None, drop_p: float = 0.0, seed: int = 0,
                seed_dev: Optional[torch.Tensor] = None, site: int = 0) -> None:
    """probs = softmax(scores) (x attention-dropout mask), dscores = softmax backward of dprobs (see the header).
    dprobs = dscores = None: forward use (softmax + dropout -> probs)."""
    assert scores.dtype == torch.float32
    out_dt = (dscores if dscores is not None else probs).dtype
    assert dprobs is None or dprobs.dtype == torch.float32 or dprobs.dtype == out_dt
    dp_is_op = dprobs is not None and dprobs.dtype != torch.float32
    ld_dp = ld_in if ld_dprobs is None else ld_dprobs
    assert kv_lens is None or kv_lens.dtype == torch.int32
    assert probs is None or probs.dtype == out_dt
    assert seed_dev is None or seed_dev.dtype == torch.int64
    lib = _lib.load()
    with _Launch("softmax_bwd" if dprobs is not None else "softmax_dropout", 8.0 * rows * n_keys + 4.0 * rows * ld_out):
        _lib.check(lib.mm_softmax_dropout_bwd(_ptr(scores), _ptr(dprobs), int(dp_is_op), ld_dp, ld_in, rows, rows_per_batch,
                                              n_keys, _ptr(kv_lens), heads, _ptr(probs), _ptr(dscores), ld_out, valid_rows,
                                              int(causal), drop_p, seed & 0xFFFFFFFFFFFFFFFF, _ptr(seed_dev), site,
                                              dtype_code(out_dt), _stream()), "mm_softmax_dropout_bwd")


def glu_bwd(pre: torch.Tensor, dy: torch.Tensor, rows: int, n: int, dpre: torch.Tensor, scale: float = 1.0) -> None:
    assert pre.dtype == dy.dtype == torch.float32 and pre.is_contiguous() and dy.is_contiguous() and dpre.is_contiguous()
    assert pre.numel() == rows * 2 * n and dy.numel() == rows * n and dpre.numel() == rows * 2 * n
    lib = _lib.load()
    with _Launch("glu_bwd", 16.0 * rows * n):
        _lib.check(lib.mm_glu_bwd(_ptr(pre), _ptr(dy), rows, n, scale, _ptr(dpre), dtype_code(dpre.dtype), _stream()),
                   "mm_glu_bwd")


def gate_bwd(z: torch.Tensor, dres_tbc: torch.Tensor, text: torch.Tensor, attn: torch.Tensor, B: int, T: int, d: int,
             dz: torch.Tensor, dcat: torch.Tensor) -> None:
    for t in (z, dres_tbc, text, attn, dcat):
        assert t.dtype == torch.float32 and t.is_contiguous()
    assert dcat.numel() == B * T * 2 * d and dz.numel() == B * T * d and dz.is_contiguous()
    lib = _lib.load()
    with _Launch("gate_bwd", 26.0 * B * T * d):
        _lib.check(lib.mm_gate_bwd(_ptr(z), _ptr(dres_tbc), _ptr(text), _ptr(attn), B, T, d, _ptr(dz), _ptr(dcat),
                                   dtype_code(dz.dtype), _stream()), "mm_gate_bwd")


def tbc_to_btc(x_tbc: torch.Tensor, B: int, T: int, d: int, out: torch.Tensor) -> None:
    assert x_tbc.dtype == out.dtype == torch.float32 and x_tbc.is_contiguous() and out.is_contiguous()
    lib = _lib.load()
    with _Launch("tbc_to_btc", 8.0 * B * T * d):
        _lib.check(lib.mm_tbc_to_btc(_ptr(x_tbc), B, T, d, _ptr(out), _stream()), "mm_tbc_to_btc")


def col2im_k5s2(dcol: torch.Tensor, B: int, t_out: int, t_in: int, C_: int, dx: torch.Tensor) -> None:
    assert dcol.dtype == dx.dtype == torch.float32 and dcol.is_contiguous() and dx.is_contiguous()
    assert dcol.numel() == B * t_out * 5 * C_ and dx.numel() == B * t_in * C_
    lib = _lib.load()
    with _Launch("col2im_k5s2", 4.0 * (dcol.numel() + dx.numel())):
        _lib.check(lib.mm_col2im_k5s2(_ptr(dcol), B, t_out, t_in, C_, _ptr(dx), _stream()), "mm_col2im_k5s2")


def grad_clip_coef(grad: torch.Tensor, grad_scale: float, max_norm: float, partials: torch.Tensor,
                   norm_coef: torch.Tensor, dev_hyper: bool = False, extra_norm: Optional[torch.Tensor] = None) -> None:
    """norm_coef[0] = ||grad_scale * grad||, norm_coef[1] = grad_scale * min(1, max_norm / (norm + 1e-6)).
    dev_hyper: grad_scale / max_norm are read from norm_coef[4] / [5] on the device (CUDA-graph replays)."""
    assert grad.dtype == torch.float32 and grad.is_contiguous() and norm_coef.numel() >= (8 if dev_hyper else 2)
    lib = _lib.load()
    assert partials.numel() >= lib.mm_sumsq_blocks()
    with _Launch("grad_clip_coef", 4.0 * grad.numel()):
        _lib.check(lib.mm_grad_clip_coef(_ptr(grad), grad.numel(), grad_scale, max_norm, _ptr(partials),
                                         _ptr(norm_coef), int(dev_hyper), _ptr(extra_norm), _stream()),
                   "mm_grad_clip_coef")
    global launch_count
    launch_count += 1   # two kernels


def adam(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, *, lr: float,
         betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, step: int = 1,
         norm_coef: Optional[torch.Tensor] = None, param_op: Optional[torch.Tensor] = None) -> None:
    """fairseq Adam step on flat fp32 buffers (in place); param_op: 16-bit copy of the new parameters (same pass)."""
    assert param_op is None or (param_op.numel() == param.numel() and param_op.is_contiguous())
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == param.numel()
    lib = _lib.load()
    with _Launch("adam", 28.0 * param.numel()):
        _lib.check(lib.mm_adam(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), lr, betas[0],
                               betas[1], eps, weight_decay, step, _ptr(norm_coef), _ptr(param_op),
                               dtype_code(param_op.dtype) if param_op is not None else 0, _stream()), "mm_adam")


def label_smoothed_nll_bwd(logits: torch.Tensor, vocab: int, target: torch.Tensor, padding_idx: int, epsilon: float,
                           dlogits: torch.Tensor, grad_scale: float = 1.0) -> None:
    """dlogits (16-bit [rows, ld_out]) = grad_scale * d label_smoothed_nll_loss(sum) / d logits."""
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
    assert target.dtype == torch.int64 and target.is_contiguous() and target.numel() == logits.shape[0]
    assert dlogits.dim() == 2 and dlogits.stride(1) == 1 and dlogits.shape[0] == logits.shape[0]
    lib = _lib.load()
    with _Launch("label_smoothed_nll_bwd", 4.0 * logits.shape[0] * vocab + 2.0 * dlogits.numel()):
        _lib.check(lib.mm_label_smoothed_nll_bwd(_ptr(logits), logits.stride(0), vocab, _ptr(target), padding_idx,
                                                 logits.shape[0], epsilon, grad_scale, _ptr(dlogits), dlogits.stride(0),
                                                 dtype_code(dlogits.dtype), _stream()), "mm_label_smoothed_nll_bwd")


def embed_tokens_bwd(tokens: torch.Tensor, padding_idx: int, dx: torch.Tensor, scale: float,
                     table_grad: torch.Tensor) -> None:
    """table_grad[v] += scale * sum of dx[row] over the rows whose token is v (deterministic; padding row skipped)."""
    assert tokens.dtype == torch.int64 and tokens.is_contiguous() and dx.dtype == table_grad.dtype == torch.float32
    assert dx.is_contiguous() and table_grad.is_contiguous()
    rows, dim = tokens.numel(), dx.shape[-1]
    lib = _lib.load()
    with _Launch("embed_tokens_bwd", 8.0 * rows * dim):
        _lib.check(lib.mm_embed_tokens_bwd(_ptr(tokens), padding_idx, _ptr(dx), rows, dim, scale, _ptr(table_grad),
                                           table_grad.numel() // dim, _stream()), "mm_embed_tokens_bwd")


def dropout(x: torch.Tensor, out: torch.Tensor, p: float, seed: int, site: int,
            resid: Optional[torch.Tensor] = None, seed_dev: Optional[torch.Tensor] = None) -> None:
    """out = (resid or 0) + dropout(x) with the counter-based mask keep(seed + seed_dev[0], site, element); in place
    allowed.  seed_dev: optional int64 device scalar (per-step seed under CUDA-graph replay)."""
    assert seed_dev is None or (seed_dev.dtype == torch.int64 and seed_dev.numel() >= 1)
    assert x.dtype == out.dtype and x.is_contiguous() and out.is_contiguous() and x.numel() == out.numel()
    assert resid is None or (resid.dtype == torch.float32 and resid.is_contiguous() and x.dtype == torch.float32)
    lib = _lib.load()
    with _Launch("dropout", float(x.element_size() * x.numel() * (2 + (resid is not None)))):
        _lib.check(lib.mm_dropout(_ptr(x), int(x.dtype == torch.float32), _ptr(resid), _ptr(out), x.numel(), p,
                                  seed & 0xFFFFFFFFFFFFFFFF, _ptr(seed_dev), site,
                                  0 if x.dtype == torch.float32 else dtype_code(x.dtype),
                                  _stream()), "mm_dropout")
