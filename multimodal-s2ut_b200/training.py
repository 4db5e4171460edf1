"""Training-step variant of the fbank -> fused-encoder path (BASELINE configs[2]): forward that keeps the
activations, backward down to every encoder parameter, gradient all-reduce, fairseq Adam.

What the reference does here is plain PyTorch autograd over the modules of the path (fairseq ``train_step``:
``loss.backward()`` -> DDP all-reduce -> ``clip_grad_norm_`` -> ``Adam.step``; scripts/textless/1_train.sh).  The
boundary of THIS path is the gradient of the fused encoder states, ``d loss / d encoder_out[0]`` ``[T, B, C]``, which
the decoder + criterion hand back ("the encoder nn.Module.forward and an autograd.Function wrapper for
training").  ``TrainEngine.backward`` turns it into the gradient of every encoder parameter:

  fusion   (mm_s2s_transformer.py:594-622, fuse.py:65-117 / :145-167)  gate -> proj -> P V -> softmax -> q k^T -> q / k|v
            projections -> image pre-norm parameters
  encoder  final LayerNorm -> L x [fc2 -> ReLU -> fc1 -> LN2 ; out_proj -> self-attention -> QKV -> LN1]
  front    x sqrt(d) -> GLU -> Conv1d #2 (col2im) -> GLU -> Conv1d #1          (fbank / CMVN have no parameters)

Every contraction is the tcgen05 GEMM of the forward pass (``mm_gemm``) with the operand layouts the backward pass
needs, so nothing is transposed or copied: dgrad = dY W reads the Linear weight as stored (MN-major W operand); wgrad =
dY^T X reads dY and X as stored (both MN-major) and splits the token contraction into the number of batches that fills
whole waves of the 74 CTA pairs (fp32 partials, summed deterministically by ``mm_reduce_partials``); attention backward
recomputes the scores per (sequence, head) straight from / into the q|k|v layout (S = q k^T, dP = dO v^T ->
``mm_softmax_bwd`` -> dV = P^T dO, dK = dS^T q, dQ = dS k).  Only the conv wgrad (strided window operand) still packs
transposed copies (``mm_pack_t``).  fp32 residual-stream gradients, 16-bit GEMM operands, fp32 parameter gradients /
Adam state in flat buffers.

Parameters and gradients live in two flat fp32 buffers (``flat_p`` / ``flat_g``; every ``nn.Parameter`` of the encoder
is re-pointed to a view), ordered so that the packed layouts the kernels produce (q|k|v, k|v, LayerNorm weight|bias)
are contiguous slices.  ``all_reduce_grads`` runs NCCL (or gloo) over ``flat_g`` in buckets; ``adam_step`` is one kernel.

Element-wise dropout (embedding, residual and activation sites, SA_image_dropout, SA_text_dropout) uses counter-based
masks that the backward pass regenerates (``mm_dropout``); attention dropout (encoder self-attention,
SA_attention_dropout) runs the un-fused score / softmax+dropout / P V path in the training forward; modality dropout
(per-batch image zeroing) works; so do image key masks (burnt into the kept scores as -inf), several image-feature
types (one attention + gate each, summed) and batches that live in the device feature store.  Not built (raises): the
audio-drop branch, which is broken in the reference (:500).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import kernels as K
from .engine import EncoderEngine, _even, _round_up, sub_len
from .feature_store import StoredImages


def _split_k(n: int, kin: int, mp: int) -> int:
    """Number of K-splits of a wgrad GEMM [n, kin] contracting over mp tokens: fill the 74 CTA pairs."""
    tiles = ((n + 255) // 256) * ((kin + 255) // 256)
    s = 1
    while s < 16 and tiles * s < 74 and (mp // 64) % (2 * s) == 0:
        s *= 2
    return s


# dropout sites (the counter-based mask is a function of (seed, site, element index))
SITE_EMBED, SITE_IMAGE, SITE_SA_ATTN, SITE_TEXT = 0, 8, 9, 10


def site_layer(i: int, which: int) -> int:
    """which: 0 = after the self-attention out_proj, 1 = after ReLU (activation dropout), 2 = after fc2,
    3 = the self-attention probabilities (attention dropout)."""
    return 16 * (i + 1) + which


def _best_split(tiles: int, M: int, pairs: int = 74, max_split: int = 32, min_chunk: int = 512) -> int:
    """Split-K factor of a wgrad GEMM with `tiles` 256 x 256 output tiles contracting over M tokens: the persistent
    kernel runs one tile per CTA pair per wave, so pick the split whose tile count fills whole waves of the 74 pairs
    best (within 3 %: the fewest partials), keeping every chunk at least `min_chunk` tokens long."""
    effs = []
    for s in range(1, max_split + 1):
        if s > 1 and M // s < min_chunk:
            break
        work = tiles * s
        effs.append((s, work / (((work + pairs - 1) // pairs) * pairs)))
    top = max(e for _, e in effs)
    return next(s for s, e in effs if e >= top - 0.03)      # fewer, longer chunks when the fill is within 3 %


class _scope:
    """Tags the launches of a phase in the per-launch timing list (``kernels.timing``)."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        self.prev, K.scope = K.scope, self.name

    def __exit__(self, *exc):
        K.scope = self.prev
        return False


def all_reduce_flat(flat: torch.Tensor, bucket_elems: int = 0, peer_memory: bool = True) -> int:
    """SUM all-reduce of a flat gradient buffer over the default process group: ONE collective by default (nothing
    overlaps it here, and on NVLink one 162 MB NCCL call takes 0.35 ms where five 32 MB buckets take 0.45:
    profiles/r02/allreduce_probe_n2.txt); bucket_elems > 0 issues asynchronous buckets of that many elements."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 1
    if bucket_elems <= 0:
        if peer_memory:
            from .peer import peer_all_reduce       # two-shot in-place kernel over NVLink peer memory (csrc/p2p.cu)

            ws = peer_all_reduce(flat)
            if ws is not None:
                return ws
        bucket_elems = max(1, flat.numel())
    works = [dist.all_reduce(flat[o:o + bucket_elems], op=dist.ReduceOp.SUM, async_op=True)
             for o in range(0, flat.numel(), bucket_elems)]
    for wk in works:
        wk.wait()
    return dist.get_world_size()


class EncoderOutGrad(torch.autograd.Function):
    """Autograd boundary of the path.  Forward: identity on the fused states.  Backward: the incoming
    ``d loss / d encoder_out`` goes through ``TrainEngine.backward`` and the parameter gradients are handed back to
    autograd as the gradients of the Function's parameter inputs -- so ``.grad`` accumulation, ``zero_grad`` in either
    mode, gradient hooks and DistributedDataParallel's reducer all behave as for any other module."""

    @staticmethod
    def forward(ctx, out, eng, generation, *params):
        ctx.eng, ctx.generation = eng, generation
        return out.view_as(out)

    @staticmethod
    def backward(ctx, grad):
        eng = ctx.eng
        if eng.generation != ctx.generation or eng._saved is None:
            # the engine keeps ONE set of saved activations (fixed workspaces): a second training forward before this
            # backward has overwritten them, and running on would give silently wrong gradients
            raise RuntimeError(
                "mm_s2ut_transformer (B200 build): backward of a training forward whose saved activations were "
                "overwritten by a later training forward of the same encoder; run forward -> backward per batch "
                "(gradient accumulation over micro-batches works: backward each micro-batch before the next forward)")
        eng.backward(grad, accumulate=False)
        g = eng.flat_g.clone()          # one 4-byte-per-parameter copy: autograd owns what it is given
        grads = []
        for p in eng.params:
            o, n = eng._slices[id(p)]
            grads.append(g[o:o + n].view(p.shape))
        return (None, None, None, *grads)


class TrainEngine(EncoderEngine):
    def __init__(self, enc, op_dtype: Optional[torch.dtype] = None, block_n: int = 256):
        self._flatten(enc)
        super().__init__(enc, op_dtype, block_n)
        # (the inference forward of this engine keeps the base class's fused GEMM+LN path)
        # training forward: fused GEMM + residual + LayerNorm with a separate output buffer (needs n == 512)
        self.train_fused_ln = self.d == 512 and getattr(enc, "fuse_layernorm", True)
        # attention backward: scores, dP and the softmax backward in one kernel (mm_attention_bwd_scores)
        self.fused_attn_bwd = getattr(enc, "fuse_attention_backward", True)
        # weight gradients of the Linear layers are queued and run pooled in one grouped launch at the end of the
        # backward pass (``mm_wgrad_grouped``: full token contraction per tile, no split-K partials); off: one split-K
        # GEMM per gradient + deferred partial reduction.  wgrad_flush_layers > 0 flushes every that many layers instead
        # (bucketed gradient all-reduce overlapping the backward pass needs complete layers early).
        self.grouped_wgrad = os.environ.get("MM_GROUPED_WGRAD", "1") != "0"
        # dV / dK / dQ of attention backward on the 128 x 64-tile kernel (mm_heads_gemm); off: mm_gemm's 256-wide tiles
        self.heads_gemm = os.environ.get("MM_HEADS_GEMM", "1") != "0"
        # sequences of up to 256 positions: the whole attention backward in one kernel (mm_attention_bwd_fused)
        self.fused_attn_bwd_onchip = os.environ.get("MM_ATTN_BWD_ONCHIP", "1") != "0"
        self.wgrad_flush_layers = int(os.environ.get("MM_WGRAD_FLUSH_LAYERS", "0"))
        # dgrad of the Linear after a LayerNorm + the LayerNorm backward + the residual add in one kernel (mm_gemm_ln_bwd).
        # Opt-in: it removes 24 launches and 1.6 GB of HBM traffic per step and is 10 us faster per call in isolation, but
        # the power-capped, graph-replayed step is not measurably faster with it
        self.fused_ln_bwd = os.environ.get("MM_FUSED_LN_BWD", "0") != "0"
        self._saved = None
        self.generation = 0            # advanced by every forward_train: ties an autograd node to ITS saved activations
        self.step_count = 0
        n = self.flat_p.numel()
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.norm_coef = torch.zeros(8, dtype=torch.float32, device=self.device)   # see mm_grad_clip_coef
        self._sumsq_partials = torch.zeros(K._lib.load().mm_sumsq_blocks(), dtype=torch.float32, device=self.device)
        self._ln_blocks = K.layernorm_bwd_blocks()
        self._pack_train()

    # ------------------------------------------------------------------------------------------
    # flat parameter / gradient buffers
    # ------------------------------------------------------------------------------------------
    def _flatten(self, enc) -> None:
        order: List[torch.nn.Parameter] = []
        seen = set()

        def add(*ps):
            for p in ps:
                if p is not None and id(p) not in seen:
                    seen.add(id(p))
                    order.append(p)

        for c in enc.subsample.conv_layers:
            add(c.weight, c.bias)
        for L in enc.transformer_layers:
            a = L.self_attn
            add(L.self_attn_layer_norm.weight, L.self_attn_layer_norm.bias)
            add(a.q_proj.weight, a.k_proj.weight, a.v_proj.weight, a.q_proj.bias, a.k_proj.bias, a.v_proj.bias)
            add(a.out_proj.weight, a.out_proj.bias, L.final_layer_norm.weight, L.final_layer_norm.bias)
            add(L.fc1.weight, L.fc1.bias, L.fc2.weight, L.fc2.bias)
        add(enc.layer_norm.weight, enc.layer_norm.bias)
        if enc.multimodal_translation_flag and enc.multimodal_attention_type is not None:
            for j in range(len(enc.gate_denses)):
                if enc.multimodal_attention_type == "selective_attention":
                    s = enc.selective_attns[j]
                    add(s.q_proj.weight, s.q_proj.bias, s.k_proj.weight, s.v_proj.weight, s.k_proj.bias, s.v_proj.bias,
                        s.proj.weight, s.proj.bias)
                else:
                    m = enc.multimodal_attns[j]
                    add(m.q_proj_weight, m.k_proj_weight, m.v_proj_weight, m.in_proj_bias, m.bias_k, m.bias_v,
                        m.out_proj.weight, m.out_proj.bias)
                add(enc.gate_denses[j].weight, enc.gate_denses[j].bias)
            pn = enc.image_pre_norm_module
            if not isinstance(pn, torch.nn.Identity):
                add(pn.weight, pn.bias)
        add(*enc.parameters())     # anything else (the reference's always-present unused projections): zero gradient
        dev = order[0].device
        if self._require_cuda and dev.type != "cuda":
            raise RuntimeError("TrainEngine needs the encoder on a CUDA device (no CPU fallback)")
        offs, total = [], 0
        for p in order:
            offs.append(total)
            total += _round_up(p.numel(), 8)    # 16-byte aligned slices of the 16-bit copy too (TMA bases)
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self._slices: Dict[int, Tuple[int, int]] = {}
        with torch.no_grad():
            for p, o in zip(order, offs):
                n = p.numel()
                self.flat_p[o:o + n].copy_(p.detach().reshape(-1).float())
                p.data = self.flat_p[o:o + n].view(p.shape)
                p.grad = self.flat_g[o:o + n].view(p.shape)      # stand-alone engine API; the module path lets autograd own .grad
                self._slices[id(p)] = (o, n)
        self.params = order
        # gradient buckets in the order the backward pass completes them: top (final LayerNorm, fusion, everything
        # else), layers L-1 .. 0, conv.  Flat layout: [conv | layer 0 .. L-1 | top]
        off = {id(p): o for p, o in zip(order, offs)}
        l_lo = [off[id(L.self_attn_layer_norm.weight)] for L in enc.transformer_layers]
        top_lo = off[id(enc.layer_norm.weight)]
        self.bucket_conv = (0, l_lo[0] if l_lo else top_lo)
        self.bucket_layers = [(lo, (l_lo[i + 1] if i + 1 < len(l_lo) else top_lo)) for i, lo in enumerate(l_lo)]
        self.bucket_top = (top_lo, total)
        self._reduce_works = []
        self._reduced = False
        self._comm_stream = None

    def g(self, *params) -> torch.Tensor:
        """Flat fp32 gradient slice spanning the given (adjacent, in this order) parameters."""
        o0, n0 = self._slices[id(params[0])]
        end = o0 + n0
        for p in params[1:]:
            o, n = self._slices[id(p)]
            assert o == end, "parameters are not adjacent in the flat buffer"
            end = o + n
        return self.flat_g[o0:end]

    def w(self, *params) -> torch.Tensor:
        o0, n0 = self._slices[id(params[0])]
        end = o0 + n0
        for p in params[1:]:
            o, n = self._slices[id(p)]
            assert o == end
            end = o + n
        return self.flat_p[o0:end]

    # ------------------------------------------------------------------------------------------
    # 16-bit operand copies for the backward pass (transposed weights, un-permuted conv weights)
    # ------------------------------------------------------------------------------------------
    def opw(self, *params) -> torch.Tensor:
        """16-bit operand copy of the given (adjacent) parameters: a slice of ``flat_op`` (written by the Adam kernel)."""
        o0, n0 = self._slices[id(params[0])]
        end = o0 + n0
        for p in params[1:]:
            o, n = self._slices[id(p)]
            assert o == end
            end = o + n
        return self.flat_op[o0:end]

    def _pack(self) -> None:
        """Forward operand set.  Linear weights are views of ``flat_op`` and biases / LayerNorm parameters views of
        ``flat_p`` (so an optimizer step needs no re-packing of them); conv weights keep their GLU tile permutation."""
        if not hasattr(self, "flat_op"):
            self.flat_op = torch.empty(self.flat_p.numel(), dtype=self.op_dtype, device=self.device)
            K.convert(self.flat_p, self.flat_op)
        if getattr(self, "_linked", False):
            self._pack_conv()
            return
        super()._pack()
        enc, d = self.enc, self.d
        for L, mod in zip(self.layers, enc.transformer_layers):
            a = mod.self_attn
            L.update(wqkv=self.opw(a.q_proj.weight, a.k_proj.weight, a.v_proj.weight).view(3 * d, d),
                     bqkv=self.w(a.q_proj.bias, a.k_proj.bias, a.v_proj.bias),
                     wo=self.opw(a.out_proj.weight).view(d, d), w1=self.opw(mod.fc1.weight).view(self.ffn, d),
                     w2=self.opw(mod.fc2.weight).view(d, self.ffn))
        for j, F in enumerate(self.fusion):
            dk = F["dk"]
            g = enc.gate_denses[j]
            if enc.multimodal_attention_type == "selective_attention":
                s = enc.selective_attns[j]
                F.update(wq=self.opw(s.q_proj.weight).view(d, d), wkv=self.opw(s.k_proj.weight, s.v_proj.weight).view(2 * d, dk),
                         bkv=self.w(s.k_proj.bias, s.v_proj.bias), wp=self.opw(s.proj.weight).view(d, d))
            else:
                m = enc.multimodal_attns[j]
                F.update(wq=self.opw(m.q_proj_weight).view(d, d), wkv=self.opw(m.k_proj_weight, m.v_proj_weight).view(2 * d, dk),
                         bq=self.w(m.in_proj_bias)[:d], bkv=self.w(m.in_proj_bias)[d:], wp=self.opw(m.out_proj.weight).view(d, d),
                         bias_kv=(self.opw(m.bias_k), self.opw(m.bias_v)))
            F.update(wg=self.opw(g.weight).view(d, 2 * d))
        self._linked = True

    def _pack_train(self) -> None:
        """Backward-only operand copies: un-permuted conv weights (the forward copies are GLU-tile permuted).  Every
        other backward GEMM reads the forward operand copies as stored (MN-major W operand for dgrad)."""
        enc = self.enc
        self.conv_bwd = []
        for i, c in enumerate(enc.subsample.conv_layers):
            cout, cin, k = c.weight.shape
            wp = self.w(c.weight).view(cout, cin, k).permute(0, 2, 1).reshape(cout, k * cin).contiguous()  # layout glue
            w_plain = self.buf(f"convw_plain{i}", (cout, k * cin), self.op_dtype)
            K.convert(wp, w_plain)
            self.conv_bwd.append(dict(w=w_plain, b=self.w(c.bias), p=c, cin=cin, cout=cout, k=k))
        self.fusion_bwd = []
        for j, F in enumerate(self.fusion):
            if enc.multimodal_attention_type == "selective_attention":
                s = enc.selective_attns[j]
                ps = dict(wq=(s.q_proj.weight,), bq=(s.q_proj.bias,), wkv=(s.k_proj.weight, s.v_proj.weight),
                          bkv=(s.k_proj.bias, s.v_proj.bias), wp=(s.proj.weight,), bp=(s.proj.bias,), bias_kv=None)
            else:
                m = enc.multimodal_attns[j]
                ps = dict(wq=(m.q_proj_weight,), wkv=(m.k_proj_weight, m.v_proj_weight), in_b=(m.in_proj_bias,),
                          wp=(m.out_proj.weight,), bp=(m.out_proj.bias,), bias_kv=(m.bias_k, m.bias_v))
            gd = enc.gate_denses[j]
            ps.update(wg=(gd.weight,), bg=(gd.bias,))
            self.fusion_bwd.append(dict(p=ps))

    def refresh_operands(self) -> None:
        """Re-derive every 16-bit operand copy from the fp32 parameters: for optimizers that update the parameters
        themselves (torch / fairseq optimizers on the module path); ``adam_step`` does it as part of its own update."""
        K.convert(self.flat_p, self.flat_op)
        self.repack()

    def repack(self) -> None:
        """Refresh every 16-bit operand copy from the fp32 master parameters (after an optimizer step)."""
        self._pack()
        self._pack_train()
        for j, F in enumerate(self.fusion):     # learned extra key / value rows live in the cached K / V^T workspaces
            if F["bias_kv"] is None:
                continue
            d = self.d
            for (name, shape, _), t in self._buf.items():
                if name == f"k{j}":                       # [B, Tk, d]
                    t[:, shape[1] - 1, :] = F["bias_kv"][0]
                elif name.startswith(f"vt_img{j}_"):      # [B, d, Tkp], one per key count Tk (in the name)
                    t[:, :, int(name.rsplit("_", 1)[1]) - 1] = F["bias_kv"][1]
                elif name == f"kv{j}":                    # [B, Tk, 2d] of the fused attention path
                    t[:, shape[1] - 1, :d] = F["bias_kv"][0]
                    t[:, shape[1] - 1, d:] = F["bias_kv"][1]

    def grads_attached(self) -> bool:
        """True when every ``param.grad`` still is this engine's view of ``flat_g`` (an optimizer's
        ``zero_grad(set_to_none=True)`` detaches them)."""
        return all(p.grad is not None and p.grad.data_ptr() == self.flat_g.data_ptr() + 4 * self._slices[id(p)][0]
                   for p in self.params)

    def attach_grads(self) -> None:
        for p in self.params:
            o, n = self._slices[id(p)]
            p.grad = self.flat_g[o:o + n].view(p.shape)

    # ------------------------------------------------------------------------------------------
    # forward (activations kept)
    # ------------------------------------------------------------------------------------------
    def _layer_train(self, i: int, x_in: torch.Tensor, B: int, T: int, seq_lens: torch.Tensor, h_next=None,
                     h_next_f32=None):
        """One pre-LN encoder layer with every sub-layer input kept.  d_model == 512: the two residual updates and the
        LayerNorms that follow them are the fused GEMM+LN kernel with a separate output (on entry ``t_h1_i`` holds
        LN1(x_in); on exit ``h_next`` holds the next LayerNorm of x_out); otherwise GEMM + stand-alone LayerNorm."""
        L, d, M, op, bn = self.layers[i], self.d, B * T, self.op_dtype, self.block_n
        s = dict(x_in=x_in, h1=self.buf(f"t_h1_{i}", (M, d), op), qkv=self.buf(f"t_qkv_{i}", (M, 3 * d), op),
                 att=self.buf(f"t_att_{i}", (M, d), op), x_mid=self.buf(f"t_xmid_{i}", (M, d), torch.float32),
                 h2=self.buf(f"t_h2_{i}", (M, d), op), f=self.buf(f"t_f_{i}", (M, self.ffn), op),
                 x_out=self.buf(f"t_xout_{i}", (M, d), torch.float32))
        p_drop, p_act, seed, seed_dev = self._drop
        fused = self.train_fused_ln      # residual-site dropout runs inside the fused GEMM + residual + LayerNorm epilogue
        dr = lambda p_, k_: (p_, seed, seed_dev, site_layer(i, k_)) if p_ > 0 else None
        if not fused:
            K.layernorm(x_in, L["ln1_g"], L["ln1_b"], out_op=s["h1"])
        K.gemm(a0=s["h1"], a0_ld=d, rows=M, w=L["wqkv"], n=3 * d, k=d, mode=K.EPI_OP, bias=L["bqkv"], scale=64 ** -0.5,
               scale_cols=d, out0=s["qkv"], out0_ld=3 * d, block_n=bn)
        if self._p_attn > 0 and K.self_attention_drop_supported(T) and self.fused_attn_bwd_onchip:
            # attention dropout generated inside the fused kernels: the forward keeps lse, the on-chip backward
            # regenerates the mask
            s["lse"] = self.buf(f"t_lse_{i}", (B, self.heads, T), torch.float32)
            s["attn_drop"] = (self._p_attn, seed, seed_dev, site_layer(i, 3))
            K.self_attention(s["qkv"], seq_lens, B, T, self.heads, s["att"], lse=s["lse"], drop=s["attn_drop"])
        elif self._p_attn > 0:
            # attention dropout: un-fused S = q k^T -> softmax + dropout -> P V on the head-mode GEMMs (the chunked
            # attention kernel has no mask generator); the backward pass regenerates the same mask
            H, Tp = self.heads, _round_up(T, 64)
            hd = dict(heads=H, head_stride=64, batches=B * H, w_batched=True, block_n=bn)
            S = self.buf("a_S", (B * H, Tp, Tp), torch.float32)
            P = self.buf("a_P", (B * H, Tp, Tp), op)
            K.gemm(a0=s["qkv"], a0_ld=3 * d, a0_bs=T * 3 * d, w=s["qkv"][:, d:], w_ld=3 * d, w_bs=T * 3 * d, out0=S,
                   rows=T, n=T, k=64, mode=K.EPI_F32, out0_ld=Tp, out0_bs=Tp * Tp, a_hm=True, w_hm=True, **hd)
            K.softmax_bwd(S, None, Tp, B * H * Tp, Tp, T, None, Tp, probs=P, kv_lens=seq_lens, heads=H, valid_rows=T,
                          drop_p=self._p_attn, seed=seed, seed_dev=seed_dev, site=site_layer(i, 3))
            K.gemm(a0=P, a0_ld=Tp, a0_bs=Tp * Tp, rows=T, k=T, w=s["qkv"][:, 2 * d:], w_ld=3 * d, w_bs=T * 3 * d,
                   w_mn=True, w_hm=True, n=64, mode=K.EPI_OP, out0=s["att"], out0_ld=d, out0_bs=T * d, out_hm=True, **hd)
        else:
            # the log-sum-exp of every score row is all the backward pass needs to rebuild the probabilities
            s["lse"] = self.buf(f"t_lse_{i}", (B, self.heads, T), torch.float32)
            K.self_attention(s["qkv"], seq_lens, B, T, self.heads, s["att"], lse=s["lse"])
        if fused:
            K.gemm_resid_ln(s["att"], L["wo"], L["bo"], x_in, L["ln2_g"], L["ln2_b"], s["h2"], x_out=s["x_mid"],
                            drop=dr(p_drop, 0))
        elif p_drop > 0:     # x_mid = x_in + dropout(out_proj(att))
            y = self.buf("t_y", (M, d), torch.float32)
            K.gemm(a0=s["att"], a0_ld=d, rows=M, w=L["wo"], n=d, k=d, mode=K.EPI_F32, bias=L["bo"], out0=y, out0_ld=d,
                   block_n=bn)
            K.dropout(y, s["x_mid"], p_drop, seed, site_layer(i, 0), resid=x_in, seed_dev=seed_dev)
            K.layernorm(s["x_mid"], L["ln2_g"], L["ln2_b"], out_op=s["h2"])
        else:
            K.gemm(a0=s["att"], a0_ld=d, rows=M, w=L["wo"], n=d, k=d, mode=K.EPI_RESID_F32, bias=L["bo"], aux0=x_in,
                   aux_ld=d, out0=s["x_mid"], out0_ld=d, block_n=bn)
            K.layernorm(s["x_mid"], L["ln2_g"], L["ln2_b"], out_op=s["h2"])
        # fc1 + ReLU + activation dropout in one epilogue
        K.gemm(a0=s["h2"], a0_ld=d, rows=M, w=L["w1"], n=self.ffn, k=d, mode=K.EPI_RELU_OP, bias=L["b1"], out0=s["f"],
               out0_ld=self.ffn, block_n=bn, drop=dr(p_act, 1))
        if fused:
            last = i + 1 == self.n_layers
            ng, nb_ = (self.ln_g, self.ln_b) if last else (self.layers[i + 1]["ln1_g"], self.layers[i + 1]["ln1_b"])
            K.gemm_resid_ln(s["f"], L["w2"], L["b2"], s["x_mid"], ng, nb_, h_next, h_next_f32, x_out=s["x_out"],
                            drop=dr(p_drop, 2))
        elif p_drop > 0:     # x_out = x_mid + dropout(fc2(f))
            y = self.buf("t_y", (M, d), torch.float32)
            K.gemm(a0=s["f"], a0_ld=self.ffn, rows=M, w=L["w2"], n=d, k=self.ffn, mode=K.EPI_F32, bias=L["b2"], out0=y,
                   out0_ld=d, block_n=bn)
            K.dropout(y, s["x_out"], p_drop, seed, site_layer(i, 2), resid=s["x_mid"], seed_dev=seed_dev)
        else:
            K.gemm(a0=s["f"], a0_ld=self.ffn, rows=M, w=L["w2"], n=d, k=self.ffn, mode=K.EPI_RESID_F32, bias=L["b2"],
                   aux0=s["x_mid"], aux_ld=d, out0=s["x_out"], out0_ld=d, block_n=bn)
        return s

    @torch.no_grad()
    def forward_train(self, src_tokens, src_lengths, imgs_list: List[torch.Tensor], img_masks_list: List,
                      drop_audio: bool = False, drop_image: bool = False, specaug=None,
                      dropout_seed: Optional[int] = None, dropout_seed_dev: Optional[torch.Tensor] = None,
                      return_all_hiddens: bool = False):
        """dropout_seed (+ the int64 device scalar dropout_seed_dev, for CUDA-graph replay) seeds the element-wise
        dropout masks of this step; default: a counter advanced per call on top of torch.initial_seed().
        return_all_hiddens: ``encoder_states`` = every layer's output [T, B, C] (the criterion of the reference always
        asks for them, criterions/speech_to_speech_criterion.py:64); they are copies outside the autograd boundary --
        gradients reach the path only through ``encoder_out``."""
        enc = self.enc
        self.generation += 1
        p_text = float(getattr(enc, "SA_text_dropout", 0.0) or 0.0)
        self._p_attn = float(getattr(enc, "attention_dropout_p", 0.0) or 0.0)
        p_sa = float(getattr(enc, "SA_attention_dropout", 0.0) or 0.0)
        p_drop, p_act = float(enc.dropout_p), float(getattr(enc, "activation_dropout_p", 0.0))
        p_img = float(getattr(enc, "SA_image_dropout", 0.0) or 0.0)
        if dropout_seed is None:
            self._drop_calls = getattr(self, "_drop_calls", 0) + 1
            dropout_seed = (torch.initial_seed() + 0x51ED270B * self._drop_calls) & 0x7FFFFFFFFFFFFFFF
        self._drop = (p_drop, p_act, int(dropout_seed), dropout_seed_dev)
        if drop_audio:
            raise NotImplementedError("audio-drop branch (broken in the reference, :500) has no backward here")
        x1, m, seq_lens, _ = self.frontend(src_tokens, src_lengths, specaug=specaug)
        B = x1.shape[0]
        x, T = self.subsample(x1, m, seq_lens)
        M, d = B * T, self.d
        if p_drop > 0:      # S2TTransformerEncoder: x = dropout_module(embed_scale * x + positions)
            K.dropout(x, x, p_drop, self._drop[2], SITE_EMBED, seed_dev=dropout_seed_dev)
        states: List[torch.Tensor] = []
        saved = dict(B=B, T=T, m=m, x1=x1, seq_lens=seq_lens, layers=[], fused=False, drop=self._drop, p_img=p_img,
                     p_attn=self._p_attn, p_sa=p_sa, p_text=0.0)
        text_f32 = self.buf("text_f32", (M, d), torch.float32)
        text_op = self.buf("text_op", (M, d), self.op_dtype)
        fused_fwd = self.train_fused_ln
        if fused_fwd:              # LN1 of layer 0 is the only stand-alone LayerNorm
            K.layernorm(x, self.layers[0]["ln1_g"], self.layers[0]["ln1_b"], out_op=self.buf("t_h1_0", (M, d), self.op_dtype))
        for i in range(self.n_layers):
            last = i + 1 == self.n_layers
            h_next = text_op if last else self.buf(f"t_h1_{i + 1}", (M, d), self.op_dtype)
            s = self._layer_train(i, x, B, T, seq_lens, h_next=h_next, h_next_f32=text_f32 if last else None)
            saved["layers"].append(s)
            x = s["x_out"]
            if return_all_hiddens:
                states.append(x.view(B, T, d).transpose(0, 1).contiguous())
        if not fused_fwd:
            K.layernorm(x, self.ln_g, self.ln_b, out_op=text_op, out_f32=text_f32)
        saved["x_final"] = x
        mask = self.pad_mask            # written by frontend()
        if imgs_list and self.fusion:
            if len(imgs_list) > len(self.fusion) or len(img_masks_list) != len(imgs_list):
                raise ValueError("imgs_list / img_masks_list do not match image_feat_dim")
            out = torch.empty(T, B, d, dtype=torch.float32, device=self.device)
            if p_text > 0:      # SA_text_dropout (:597): the dropped states feed the query, the gate and the mix
                K.dropout(text_f32, text_f32, p_text, self._drop[2], SITE_TEXT, seed_dev=dropout_seed_dev)
                K.convert(text_f32, text_op)
                saved["p_text"] = p_text
            imgs = []
            # one attention + gate per image-feature type, fused states = sum over the types (reference :513-530, :557-560)
            for j, (img, img_mask) in enumerate(zip(imgs_list, img_masks_list)):
                if isinstance(img, StoredImages):
                    # a batch that lives in the device feature store: the backward pass (image pre-norm parameter
                    # gradients) needs the rows as a tensor -> gather them once (layout glue, like the reference's collater)
                    img = img.store.data.index_select(0, img.index.to(self.device))
                if drop_image:      # modality dropout: every image tensor zeroed (:504-505)
                    img = torch.zeros(tuple(img.shape), dtype=torch.float32, device=self.device)
                img = img.to(self.device, non_blocking=True).float().contiguous()
                res = out if j == 0 else self.buf(f"res_tbc{j}", (T, B, d), torch.float32)
                self.fuse(j, text_f32, text_op, img, img_mask, B, T, res,
                          img_dropout=(p_img, self._drop[2], dropout_seed_dev, SITE_IMAGE + 32 * j),
                          attn_dropout=(p_sa, self._drop[2], dropout_seed_dev, SITE_SA_ATTN + 32 * j), keep_scores=True,
                          tag=str(j))
                if j > 0:
                    K.dropout(res, out, 0.0, 0, 0, resid=out)      # out += res (p = 0: the plain fp32 accumulate)
                imgs.append(img)
            saved.update(fused=True, imgs=imgs)
        else:
            out = text_f32.view(B, T, d).transpose(0, 1).contiguous()
        self._saved = saved
        return {"encoder_out": [out], "encoder_padding_mask": [mask], "encoder_embedding": [], "encoder_states": states,
                "src_tokens": [], "src_lengths": []}

    # ------------------------------------------------------------------------------------------
    # backward building blocks
    # ------------------------------------------------------------------------------------------
    def _mp(self, M: int) -> int:
        return _round_up(M, 1024)

    _ARENA = 64 * 1024 * 1024      # floats: partial sums of the reductions a layer defers (256 MB)

    def _partials(self, n: int) -> torch.Tensor:
        """n floats of partial-sum workspace that stay untouched until the next ``_flush()`` (bump allocation)."""
        if n > self._ARENA:
            self._flush()
            return self.buf(f"wgrad_part_{n}", (n,), torch.float32)
        arena = self.buf("reduce_arena", (self._ARENA,), torch.float32)
        off = getattr(self, "_arena_off", 0)
        if off + n > self._ARENA:
            self._flush()
            off = 0
        self._arena_off = off + _round_up(n, 64)
        return arena[off:off + n]

    def _defer(self, part: torch.Tensor, S: int, stride: int, n: int, out: torch.Tensor, accumulate: bool) -> None:
        """Queue out[:n] (+)= sum_s part[s * stride : s * stride + n]; runs at the next ``_flush()``."""
        if not hasattr(self, "_jobs"):
            self._jobs = []
        self._jobs.append((part, S, stride, n, out, accumulate))

    def _flush(self) -> None:
        """Run the queued reductions in one launch (per 16) and release the partial-sum workspace."""
        jobs = getattr(self, "_jobs", None)
        if jobs:
            K.reduce_partials_many(jobs)
            self._jobs = []
        self._arena_off = 0

    def _lnp(self) -> torch.Tensor:
        """Per-block (dgamma, dbeta) partial workspace of one LayerNorm backward."""
        return self._partials(self._ln_blocks * 2 * 1024)

    def _wgrad(self, dyt: torch.Tensor, xt: torch.Tensor, n: int, kin: int, mp: int, out: torch.Tensor,
               accumulate: bool) -> None:
        """out [n, kin] (+)= dY^T X from TRANSPOSED token-contiguous copies dyt [n, mp], xt [kin, mp] (zero beyond M).
        Only the conv layers use this form (their X operand is a strided window view per utterance)."""
        S = _split_k(n, kin, mp)
        chunk = mp // S
        part = self._partials(S * n * kin)
        K.gemm(a0=dyt, a0_ld=dyt.shape[-1], a0_bs=chunk, rows=n, batches=S, w=xt, w_ld=xt.shape[-1], w_bs=chunk,
               w_batched=True, n=kin, k=chunk, mode=K.EPI_F32, out0=part, out0_ld=kin, out0_bs=n * kin,
               block_n=self.block_n)
        K.reduce_partials(part, S, n * kin, n * kin, out, accumulate)

    def _wgrad_mn(self, dy_op: torch.Tensor, dy_ld: int, x_ops, M: int, n: int, gw: torch.Tensor, accumulate: bool,
                  gb: Optional[torch.Tensor] = None) -> None:
        """gw [n, sum kin] (+)= dy^T [x_0 | x_1 | ...]: both operands read as stored (MN-major), the token contraction
        split into S batches; x_ops = [(tensor [M, kin], ld, kin), ...] fill consecutive column blocks of gw."""
        kin_all = sum(x[2] for x in x_ops)
        if getattr(self, "grouped_wgrad", False):
            # queued: the operands must stay untouched until _wgrad_flush() (the backward pass gives every layer its own
            # gradient buffers in this mode)
            q = self.__dict__.setdefault("_wq", {}).setdefault(self._wq_key(M, accumulate), [])
            col = 0
            for x, x_ld, kin in x_ops:     # gb: the bias gradient (column sums of dy) rides along with the first operand
                q.append((dy_op, dy_ld, x, x_ld, gw.view(-1)[col:], kin_all, n, kin, gb if col == 0 else None, M))
                col += kin
            return
        assert gb is None
        tiles = ((n + 255) // 256) * sum((x[2] + 255) // 256 for x in x_ops)
        S = _best_split(tiles, M)
        chunk = _round_up((M + S - 1) // S, 64)
        part = self._partials(S * n * kin_all)
        col = 0
        for x, x_ld, kin in x_ops:
            K.gemm(a0=dy_op, a0_ld=dy_ld, rows=n, batches=S, w=x, w_ld=x_ld, w_batched=True, n=kin, k=chunk,
                   mode=K.EPI_F32, out0=part[col:], out0_ld=kin_all, out0_bs=n * kin_all, block_n=self.block_n,
                   a_mn=True, w_mn=True, a_kbatch=True, w_kbatch=True, a_k_total=M, w_k_total=M)
            col += kin
        self._defer(part, S, n * kin_all, n * kin_all, gw, accumulate)

    @staticmethod
    def _wq_key(M: int, accumulate: bool):
        """Pool of a queued weight gradient: one pool for all token counts (the library orders the tiles of unequal groups
        by cost); MM_WGRAD_MERGE_TOKENS=0 keeps one pool -- one launch -- per token count."""
        return (0 if os.environ.get("MM_WGRAD_MERGE_TOKENS", "1") != "0" else M, bool(accumulate))

    def _wgrad_flush(self) -> None:
        """Run the queued weight gradients: one grouped launch per pool (every group carries its own token count)."""
        for (_, acc), q in self.__dict__.get("_wq", {}).items():
            if q:
                K.wgrad_grouped(q, q[0][9], acc)
                q.clear()

    def _bias_grad(self, dy_op: torch.Tensor, dy_ld: int, M: int, n: int, gb: torch.Tensor, accumulate: bool,
                   period: int = 0, valid: int = 0) -> None:
        if getattr(self, "grouped_wgrad", False) and period == 0:
            q = self.__dict__.setdefault("_wq", {}).setdefault(self._wq_key(M, accumulate), [])
            q.append((dy_op, dy_ld, None, 0, None, 0, n, 0, gb, M))
            return
        part = self._partials(K.colsum_blocks(M) * n)
        nb = K.colsum(dy_op, dy_ld, M, n, part, period, valid)
        self._defer(part, nb, n, n, gb, accumulate)

    def _linear_bwd(self, dy_op: torch.Tensor, dy_ld: int, x_op: torch.Tensor, M: int, n: int, kin: int,
                    gw: torch.Tensor, gb: Optional[torch.Tensor], accumulate: bool) -> None:
        """Parameter gradients of y = x W^T + b from the 16-bit dy [M, n] and x [M, kin]."""
        if getattr(self, "grouped_wgrad", False):
            self._wgrad_mn(dy_op, dy_ld, [(x_op, x_op.stride(0), kin)], M, n, gw, accumulate, gb)
            return
        self._wgrad_mn(dy_op, dy_ld, [(x_op, x_op.stride(0), kin)], M, n, gw, accumulate)
        if gb is not None:
            self._bias_grad(dy_op, dy_ld, M, n, gb, accumulate)

    def _dgrad_ln_bwd(self, dy_op: torch.Tensor, w_op: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor,
                      g: torch.Tensor, g_op: torch.Tensor, gwb: torch.Tensor, accumulate: bool, drop) -> None:
        """Backward through ``Linear(LayerNorm(x))`` down to the residual stream: dh = dy W, g += LayerNorm'(dh), g_op =
        16-bit g (x the dropout mask of the branch that reads it), LayerNorm parameter gradients queued.  One kernel at
        d_model 512 (``mm_gemm_ln_bwd``: dh never leaves TMEM), else the dgrad GEMM + the LayerNorm-backward row kernel."""
        M, k = dy_op.shape
        d = self.d
        if getattr(self, "fused_ln_bwd", False) and d == 512:
            part = self._partials(K.gemm_ln_bwd_partial_rows(M) * 2 * d)
            nb = K.gemm_ln_bwd(dy_op, w_op, x, gamma, g, g_op, part, drop=drop)
            self._defer(part, nb, 2 * d, 2 * d, gwb, accumulate)
            return
        dh = self.buf("b_dh", (M, d), torch.float32)
        K.gemm(a0=dy_op, a0_ld=dy_op.stride(0), rows=M, w=w_op, w_ld=d, w_mn=True, n=d, k=k, mode=K.EPI_F32, out0=dh,
               out0_ld=d, block_n=self.block_n)
        lnp = self._lnp()
        K.layernorm_bwd(x, gamma, dh, lnp, dx=g, resid=g, dx_op=g_op, drop=drop)
        self._ln_param_grads(lnp, d, gwb, accumulate)

    def _ln_param_grads(self, part: torch.Tensor, dim: int, gwb: torch.Tensor, accumulate: bool) -> None:
        self._defer(part, self._ln_blocks, 2 * dim, 2 * dim, gwb, accumulate)

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def _attention_bwd(self, s: dict, datt: torch.Tensor, dqkv: torch.Tensor, B: int, T: int,
                       seq_lens: torch.Tensor, layer_index: int = 0) -> None:
        """dqkv [M, 3d] (16-bit; q part already x head_dim^-0.5) from datt [M, d] and the saved q|k|v.  All five
        contractions read q|k|v / datt and write dqkv in place of their (sequence, head) column blocks."""
        d, H, op, bn = self.d, self.heads, self.op_dtype, self.block_n
        Tp = _round_up(T, 64)
        qkv = s["qkv"]
        if "lse" in s and self.fused_attn_bwd and self.fused_attn_bwd_onchip and T <= 256:
            # one kernel: S, dP, P, dS never leave the SM; dq | dk | dv accumulate in TMEM
            K.attention_bwd_fused(qkv, T, seq_lens, B, H, datt, s["att"], s["lse"], dqkv, drop=s.get("attn_drop"))
            return
        if "lse" in s and self.fused_attn_bwd and self.fused_attn_bwd_onchip:
            # longer sequences: the same on chip with query tiles taken in pairs (mm_attention_bwd_general)
            scratch = self.buf("a_bwd_scratch", (K.attention_bwd_general_scratch_floats(T),), torch.float32)
            K.attention_bwd_general(qkv, T, qkv[:, d:], qkv[:, 2 * d:], T, seq_lens, B, H, datt, s["att"], s["lse"], dqkv,
                                    dqkv[:, d:], dqkv[:, 2 * d:], scratch, drop=s.get("attn_drop"))
            return
        BH = B * H
        hd = dict(heads=H, head_stride=64, batches=BH, w_batched=True, block_n=bn)
        P = self.buf("a_P", (BH, Tp, Tp), op)
        dS = self.buf("a_dS", (BH, Tp, Tp), op)
        if "lse" in s and self.fused_attn_bwd:
            # one kernel: S = q k^T and dP = dO v^T in TMEM, P = exp(S - lse), dS = P o (dP - rowsum(dO o O))
            K.attention_bwd_scores(qkv, 0, T, qkv, d, qkv, 2 * d, T, seq_lens, B, H, datt, s["att"], s["lse"], P, dS)
        else:   # attention dropout (the mask is regenerated by the row kernel): scores and dP through HBM
            S = self.buf("a_S", (BH, Tp, Tp), torch.float32)
            dP = self.buf("a_dP16", (BH, Tp, Tp), op)      # 16-bit: it is a gradient (dS is 16-bit anyway); S stays fp32
            sc = dict(rows=T, n=T, k=64, out0_ld=Tp, out0_bs=Tp * Tp, a_hm=True, w_hm=True, **hd)
            K.gemm(a0=qkv, a0_ld=3 * d, a0_bs=T * 3 * d, w=qkv[:, d:], w_ld=3 * d, w_bs=T * 3 * d, out0=S, mode=K.EPI_F32, **sc)
            K.gemm(a0=datt, a0_ld=d, a0_bs=T * d, w=qkv[:, 2 * d:], w_ld=3 * d, w_bs=T * 3 * d, out0=dP, mode=K.EPI_OP, **sc)
            _, _, seed, seed_dev = self._saved["drop"]
            K.softmax_bwd(S, dP, Tp, BH * Tp, Tp, T, dS, Tp, probs=P, kv_lens=seq_lens, heads=H, valid_rows=T,
                          drop_p=self._saved["p_attn"], seed=seed, seed_dev=seed_dev, site=site_layer(layer_index, 3))
        if self.heads_gemm:     # 128 x 64 tiles per (sequence, head): no wasted columns
            hg = dict(a_ld=Tp, a_bs=Tp * Tp, out_ld=3 * d, out_bs=T * 3 * d, rows=T, k=T, batch=B, heads=H)
            K.heads_gemm(P, transposed=True, w=datt, w_ld=d, w_bs=T * d, out=dqkv[:, 2 * d:], **hg)             # dV = P^T dO
            K.heads_gemm(dS, transposed=True, w=qkv, w_ld=3 * d, w_bs=T * 3 * d, out=dqkv[:, d:], **hg)         # dK = dS^T q
            K.heads_gemm(dS, transposed=False, w=qkv[:, d:], w_ld=3 * d, w_bs=T * 3 * d, out=dqkv,               # dQ = dS k
                         scale=64 ** -0.5, **hg)
            return
        og = dict(rows=T, n=64, k=T, mode=K.EPI_OP, out0_ld=3 * d, out0_bs=T * 3 * d, out_hm=True, w_mn=True, w_hm=True,
                  a0_ld=Tp, a0_bs=Tp * Tp, **hd)
        K.gemm(a0=P, a_mn=True, w=datt, w_ld=d, w_bs=T * d, out0=dqkv[:, 2 * d:], **og)          # dV = P^T dO
        K.gemm(a0=dS, a_mn=True, w=qkv, w_ld=3 * d, w_bs=T * 3 * d, out0=dqkv[:, d:], **og)      # dK = dS^T q
        K.gemm(a0=dS, w=qkv[:, d:], w_ld=3 * d, w_bs=T * 3 * d, out0=dqkv, scale=64 ** -0.5,     # dQ = dS k
               scale_cols=64, **og)

    def _layer_bwd(self, i: int, g: torch.Tensor, g_op: torch.Tensor, B: int, T: int, seq_lens: torch.Tensor,
                   accumulate: bool) -> torch.Tensor:
        """g = d loss / d x_out [M, d] fp32 with its 16-bit copy g_op; g is overwritten with d loss / d x_in, whose 16-bit
        copy is returned (g_op itself unless the weight gradients are queued: their operands -- g_op, dF, dqkv of every
        layer -- then live in per-layer buffers until the grouped launch)."""
        tag = f"@{i}" if self.grouped_wgrad else ""
        s, L = self._saved["layers"][i], self.layers[i]
        mod = self.enc.transformer_layers[i]
        a = mod.self_attn
        d, ffn, M, op, bn = self.d, self.ffn, B * T, self.op_dtype, self.block_n
        p_drop, p_act, seed, seed_dev = self._saved["drop"]
        # with dropout the 16-bit gradient copy arrives ALREADY masked for this layer's fc2 branch (gradient entering fc2 =
        # g o mask / (1 - p); the residual branch keeps the fp32 g): the LayerNorm backward that produced it burnt the
        # mask of site (i, 2) in, and the one below burns in the mask of site (i, 0) for the attention branch
        gm = g_op
        drb = lambda k_, li=i: (p_drop, seed, seed_dev, site_layer(li, k_)) if p_drop > 0 else None
        # ---- FFN: x_out = x_mid + dropout(fc2(dropout(relu(fc1(LN2(x_mid))))))
        self._linear_bwd(gm, d, s["f"], M, d, ffn, self.g(mod.fc2.weight), self.g(mod.fc2.bias), accumulate)
        dF = self.buf("b_dF" + tag, (M, ffn), op)
        # ReLU (and activation-dropout) mask in the dgrad's epilogue: the kept activation is > 0 exactly where ReLU passed
        # and the dropout kept it; the surviving gradient is scaled by 1 / (1 - p_act)
        K.gemm(a0=gm, a0_ld=d, rows=M, w=L["w2"], w_ld=ffn, w_mn=True, n=ffn, k=d, mode=K.EPI_MASK_OP, out0=dF,
               out0_ld=ffn, aux0=s["f"], aux_ld=ffn, scale=1.0 / (1.0 - p_act), block_n=bn)
        self._linear_bwd(dF, ffn, s["h2"], M, ffn, d, self.g(mod.fc1.weight), self.g(mod.fc1.bias), accumulate)
        if tag:
            g_op = self.buf("b_g_op_mid" + tag, (M, d), op)
        self._dgrad_ln_bwd(dF, L["w1"], s["x_mid"], L["ln2_g"], g, g_op,
                           self.g(mod.final_layer_norm.weight, mod.final_layer_norm.bias), accumulate, drb(0))
        # ---- attention: x_mid = x_in + dropout(out_proj(attn(LN1(x_in))))
        gm = g_op
        self._linear_bwd(gm, d, s["att"], M, d, d, self.g(a.out_proj.weight), self.g(a.out_proj.bias), accumulate)
        datt = self.buf("b_datt", (M, d), op)
        K.gemm(a0=gm, a0_ld=d, rows=M, w=L["wo"], w_ld=d, w_mn=True, n=d, k=d, mode=K.EPI_OP, out0=datt, out0_ld=d,
               block_n=bn)
        dqkv = self.buf("b_dqkv" + tag, (M, 3 * d), op)
        with _scope("attn"):
            self._attention_bwd(s, datt, dqkv, B, T, seq_lens, i)
        self._linear_bwd(dqkv, 3 * d, s["h1"], M, 3 * d, d, self.g(a.q_proj.weight, a.k_proj.weight, a.v_proj.weight),
                         self.g(a.q_proj.bias, a.k_proj.bias, a.v_proj.bias), accumulate)
        if tag:
            g_op = self.buf("b_g_op_in" + tag, (M, d), op)
        # the copy that leaves this layer enters the fc2 branch of layer i - 1 (nothing below layer 0 reads it)
        self._dgrad_ln_bwd(dqkv, L["wqkv"], s["x_in"], L["ln1_g"], g, g_op,
                           self.g(mod.self_attn_layer_norm.weight, mod.self_attn_layer_norm.bias), accumulate,
                           drb(2, i - 1) if i > 0 else None)
        if not self.grouped_wgrad:      # split-K partials fill the arena: reduce per layer.  Pooled weight gradients leave
            self._flush()               # only the LayerNorm partials (4.8 MB per layer): reduced with the pool's flush
        return g_op

    def _fusion_bwd(self, j: int, dres: torch.Tensor, gtext: torch.Tensor, B: int, T: int, accumulate: bool,
                    ln_accumulate: bool) -> None:
        """Image type j: dres [T, B, d] -> gtext [M, d] = d loss / d text (the final LayerNorm output) through this
        type's attention + gate, and the type's parameter gradients.  ln_accumulate: the shared image pre-norm's
        gradient already holds an earlier type's contribution."""
        enc, F, Fb = self.enc, self.fusion[j], self.fusion_bwd[j]
        ps = Fb["p"]
        d, M, op, bn, dk = self.d, B * T, self.op_dtype, self.block_n, F["dk"]
        img = self._saved["imgs"][j]
        tag = str(j)
        Tk_img = img.shape[1]
        extra = 1 if F["bias_kv"] is not None else 0
        Tk = Tk_img + extra
        Tkp = _round_up(Tk, 8)
        text_f32, text_op = self.buf("text_f32", (M, d), torch.float32), self.buf("text_op", (M, d), op)
        o = self.buf("o_img" + tag, (M, d), op)
        q = self.buf("q_img" + tag, (M, d), op)
        kbuf = self.buf(f"k{j}", (B, Tk, d), op)
        vt = self.buf(f"vt_img{j}_{Tk}", (B, d, Tkp), op)
        S = self.buf(f"S{j}", (B, T, Tkp), torch.float32)
        P = self.buf(f"P{j}", (B, T, Tkp), op)
        img_op = self.buf(f"img_op{j}", (B * Tk_img, dk), op)
        da_op = self.buf("f_da_op" + tag, (M, d), op)
        if enc.use_selective_gate:
            a_f32, a_op = self.buf("attn_f32" + tag, (M, d), torch.float32), self.buf("attn_op" + tag, (M, d), op)
            z = self.buf("f_z", (M, d), torch.float32)
            K.gemm(a0=a_op, a0_ld=d, a1=text_op, a1_ld=d, k_split=d, rows=M, w=F["wg"], n=d, k=2 * d, mode=K.EPI_F32,
                   bias=F["bg"], out0=z, out0_ld=d, block_n=bn)
            dz = self.buf("f_dz" + tag, (M, d), op)
            dcat = self.buf("f_dcat" + tag, (M, 2 * d), torch.float32)
            K.gate_bwd(z, dres, text_f32, a_f32, B, T, d, dz, dcat)
            # gate Linear(2d -> d) on [attn | text]: the two operand halves fill the two column blocks of dWg
            self._wgrad_mn(dz, d, [(a_op, d, d), (text_op, d, d)], M, d, self.g(*ps["wg"]), accumulate)
            self._bias_grad(dz, d, M, d, self.g(*ps["bg"]), accumulate)
            K.gemm(a0=dz, a0_ld=d, rows=M, w=F["wg"], w_ld=2 * d, w_mn=True, n=2 * d, k=d, mode=K.EPI_RESID_F32,
                   aux0=dcat, aux_ld=2 * d, out0=dcat, out0_ld=2 * d, block_n=bn)
            da, da_ld, dtext_part, dtext_ld = dcat, 2 * d, dcat[:, d:], 2 * d
        else:   # res = text + attn
            dflat = self.buf("f_dflat" + tag, (M, d), torch.float32)
            K.tbc_to_btc(dres, B, T, d, dflat)
            da, da_ld, dtext_part, dtext_ld = dflat, d, dflat, d
        # ---- proj: attn = o Wp^T + bp
        K.pack_t(da, rows=M, cols=d, in_ld=da_ld, out_n=da_op, n_ld=d)
        self._linear_bwd(da_op, d, o, M, d, d, self.g(*ps["wp"]), self.g(*ps["bp"]), accumulate)
        do = self.buf("f_do", (M, d), op)
        K.gemm(a0=da_op, a0_ld=d, rows=M, w=F["wp"], w_ld=d, w_mn=True, n=d, k=d, mode=K.EPI_OP, out0=do, out0_ld=d,
               block_n=bn)
        # ---- o = P V ; P = softmax(q k^T): V^T, P, dS, q, k are all read as stored
        bt = dict(batches=B, w_batched=True, block_n=bn)
        dP = self.buf("f_dP", (B, T, Tkp), torch.float32)
        K.gemm(a0=do, a0_ld=d, a0_bs=T * d, rows=T, w=vt, w_ld=Tkp, w_bs=d * Tkp, w_mn=True, n=Tk, k=d, mode=K.EPI_F32,
               out0=dP, out0_ld=Tkp, out0_bs=T * Tkp, **bt)                                        # dP = dO V^T
        dS = self.buf("f_dS", (B, T, Tkp), op)
        K.softmax_bwd(S, dP, Tkp, M, T, Tk, dS, Tkp, drop_p=self._saved["p_sa"], seed=self._saved["drop"][2],
                      seed_dev=self._saved["drop"][3], site=SITE_SA_ATTN + 32 * j)      # (a key mask lives in S as -inf)
        dkv = self.buf("f_dkv" + tag, (B, Tk, 2 * d), op)
        kvg = dict(rows=Tk, a0_ld=Tkp, a0_bs=T * Tkp, a_mn=True, w_ld=d, w_bs=T * d, w_mn=True, n=d, k=T, mode=K.EPI_OP,
                   out0_ld=2 * d, out0_bs=Tk * 2 * d, **bt)
        K.gemm(a0=dS, w=q, out0=dkv, **kvg)                                                       # dK = dS^T q
        K.gemm(a0=P, w=do, out0=dkv.view(-1)[d:], **kvg)                                          # dV = P^T dO
        dq = self.buf("f_dq" + tag, (M, d), op)
        K.gemm(a0=dS, a0_ld=Tkp, a0_bs=T * Tkp, rows=T, w=kbuf, w_ld=d, w_bs=Tk * d, w_mn=True, n=d, k=Tk, mode=K.EPI_OP,
               scale=d ** -0.5, scale_cols=d, out0=dq, out0_ld=d, out0_bs=T * d, **bt)            # dq = dS k
        # ---- q projection: parameter grads + text gradient
        gbq = self.g(*ps["bq"]) if extra == 0 else self.g(*ps["in_b"])[:d]
        self._linear_bwd(dq, d, text_op, M, d, d, self.g(*ps["wq"]), gbq, accumulate)
        K.gemm(a0=dq, a0_ld=d, rows=M, w=F["wq"], w_ld=d, w_mn=True, n=d, k=d, mode=K.EPI_RESID_F32, aux0=dtext_part,
               aux_ld=dtext_ld, out0=gtext, out0_ld=d, block_n=bn)
        # ---- k|v projection over the image tokens
        gbkv = self.g(*ps["bkv"]) if extra == 0 else self.g(*ps["in_b"])[d:]
        if extra == 0:
            self._linear_bwd(dkv, 2 * d, img_op, B * Tk_img, 2 * d, dk, self.g(*ps["wkv"]), gbkv, accumulate)
        else:
            # the learned extra key / value row of every utterance is not a projected image token: contract per utterance
            # over its Tk_img image rows (batched MN-major operands), then sum the per-utterance partials
            part = self._partials(B * 2 * d * dk)
            K.gemm(a0=dkv, a0_ld=2 * d, a0_bs=Tk * 2 * d, a_mn=True, rows=2 * d, w=img_op, w_ld=dk, w_bs=Tk_img * dk,
                   w_mn=True, n=dk, k=Tk_img, mode=K.EPI_F32, out0=part, out0_ld=dk, out0_bs=2 * d * dk, **bt)
            self._defer(part, B, 2 * d * dk, 2 * d * dk, self.g(*ps["wkv"]), accumulate)
            self._bias_grad(dkv, 2 * d, B * Tk, 2 * d, gbkv, accumulate, period=Tk, valid=Tk_img)
            # learned bias_k | bias_v: key / value number Tk_img of every utterance
            self._bias_grad(dkv.view(-1)[Tk_img * 2 * d:], Tk * 2 * d, B, 2 * d, self.g(*ps["bias_kv"]), accumulate)
        if self.img_ln is not None:
            dimg = self.buf("f_dimg", (B * Tk_img, dk), torch.float32)
            K.gemm(a0=dkv, a0_ld=2 * d, a0_bs=Tk * 2 * d, rows=Tk_img, batches=B, w=F["wkv"], w_ld=dk, w_mn=True, n=dk,
                   k=2 * d, mode=K.EPI_F32, out0=dimg, out0_ld=dk, out0_bs=Tk_img * dk, block_n=bn)
            if self._saved["p_img"] > 0:     # SA_image_dropout sat between the pre-norm and the K|V projection
                _, _, seed, seed_dev = self._saved["drop"]
                K.dropout(dimg, dimg, self._saved["p_img"], seed, SITE_IMAGE + 32 * j, seed_dev=seed_dev)
            lnp = self._lnp()
            K.layernorm_bwd(img.view(B * Tk_img, dk), self.img_ln[0], dimg, lnp)
            pn = enc.image_pre_norm_module
            self._ln_param_grads(lnp, dk, self.g(pn.weight, pn.bias), accumulate or ln_accumulate)

    def _conv_bwd(self, g: torch.Tensor, B: int, T: int, accumulate: bool) -> None:
        """g = d loss / d x0 [B*T, d] (x0 = glu(conv2) * sqrt(d) + positions) -> conv parameter gradients."""
        sv, op, bn, d = self._saved, self.op_dtype, self.block_n, self.d
        c1, c2 = self.conv_bwd
        m, x1 = sv["m"], sv["x1"]
        m_alloc = x1.shape[1]
        T1, mid = sub_len(m), c1["cout"] // 2
        T1_alloc = _even(T1 + 4)
        x2 = self.buf("x2", (B, T1_alloc, mid), op, zero=True)

        def conv_param_grads(c, dpre, xin, in_ld, in_bs, Tout, kin):
            Te = _even(Tout)
            mp = self._mp(B * Te)
            n = c["cout"]
            dyt = self.buf(f"tc_{n}_{mp}", (n, mp), op, zero=True)
            xt = self.buf(f"tcx_{kin}_{mp}", (kin, mp), op, zero=True)
            K.pack_t(dpre, rows=Tout, cols=n, in_ld=n, batches=B, in_bs0=Tout * n, out_t=dyt, t_ld=mp, t_bs0=Te,
                     t_cols_pad=Te)
            K.pack_t(xin, rows=Tout, cols=kin, in_ld=in_ld, batches=B, in_bs0=in_bs, out_t=xt, t_ld=mp, t_bs0=Te,
                     t_cols_pad=Te)
            gw = self.buf(f"convgw_{n}_{kin}", (n, kin), torch.float32)
            self._wgrad(dyt, xt, n, kin, mp, gw, False)
            # layout glue: [cout, tap, cin] (GEMM operand order) -> Conv1d.weight [cout, cin, tap]
            gview = gw.view(n, c["k"], c["cin"]).permute(0, 2, 1)
            gflat = self.g(c["p"].weight).view(n, c["cin"], c["k"])
            if accumulate:
                gflat.add_(gview)
            else:
                gflat.copy_(gview)
            K.rowsum(dyt, mp, n, B * Te, self.g(c["p"].bias), accumulate)

        # ---- conv 2 + GLU (+ x sqrt(d))
        M = B * T
        pre2 = self.buf("c_pre2", (M, c2["cout"]), torch.float32)
        K.gemm(a0=x2, a0_ld=2 * mid, a0_bs=T1_alloc * mid, rows=T, batches=B, w=c2["w"], n=c2["cout"], k=c2["k"] * mid,
               mode=K.EPI_F32, bias=c2["b"], out0=pre2, out0_ld=c2["cout"], out0_bs=T * c2["cout"], block_n=bn)
        dpre2 = self.buf("c_dpre2", (M, c2["cout"]), op)
        K.glu_bwd(pre2, g, M, d, dpre2, scale=self.embed_scale)
        conv_param_grads(c2, dpre2, x2, 2 * mid, T1_alloc * mid, T, c2["k"] * mid)
        dcol = self.buf("c_dcol", (M, c2["k"] * mid), torch.float32)
        K.gemm(a0=dpre2, a0_ld=c2["cout"], rows=M, w=c2["w"], w_ld=c2["k"] * mid, w_mn=True, n=c2["k"] * mid,
               k=c2["cout"], mode=K.EPI_F32, out0=dcol, out0_ld=c2["k"] * mid, block_n=bn)
        dglu1 = self.buf("c_dglu1", (B * T1, mid), torch.float32)
        K.col2im_k5s2(dcol, B, T, T1, mid, dglu1)
        # ---- conv 1 + GLU
        M1 = B * T1
        pre1 = self.buf("c_pre1", (M1, c1["cout"]), torch.float32)
        K.gemm(a0=x1, a0_ld=2 * c1["cin"], a0_bs=m_alloc * c1["cin"], rows=T1, batches=B, w=c1["w"], n=c1["cout"],
               k=c1["k"] * c1["cin"], mode=K.EPI_F32, bias=c1["b"], out0=pre1, out0_ld=c1["cout"],
               out0_bs=T1 * c1["cout"], block_n=bn)
        dpre1 = self.buf("c_dpre1", (M1, c1["cout"]), op)
        K.glu_bwd(pre1, dglu1, M1, mid, dpre1)
        conv_param_grads(c1, dpre1, x1, 2 * c1["cin"], m_alloc * c1["cin"], T1, c1["k"] * c1["cin"])

    def _reduce_async(self, lo: int, hi: int) -> None:
        """All-reduce flat_g[lo:hi] over the ranks while the backward pass continues: the collective is issued on a side
        stream behind an event that marks the bucket complete (NCCL over NVLink; plain blocking call under gloo)."""
        import torch.distributed as dist

        if hi <= lo:
            return
        bucket = self.flat_g[lo:hi]
        if self.device.type != "cuda":
            dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        from .peer import peer_group

        grp = peer_group(self.flat_g)      # created on first use (outside graph capture: the eager warm-up pass)
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ev)
            if grp is not None and lo % grp.align == 0 and (hi % grp.align == 0 or hi == self.flat_g.numel()):
                grp.all_reduce(lo, hi)      # barrier + exchange + barrier kernels: captured with the backward pass
            else:
                self._reduce_works.append(dist.all_reduce(bucket, op=dist.ReduceOp.SUM, async_op=True))

    def _reduce_join(self) -> None:
        for wk in self._reduce_works:
            wk.wait()
        self._reduce_works = []
        if self._comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
        self._reduced = True

    @torch.no_grad()
    def backward(self, grad_out: torch.Tensor, accumulate: bool = False, overlap_reduce: bool = False) -> None:
        """grad_out = d loss / d encoder_out[0]  [T, B, d] fp32.  Fills (or accumulates into) every ``param.grad``.
        overlap_reduce: all-reduce each gradient bucket (top, every layer, conv) over the process group as soon as the
        backward pass has completed it, overlapping the collective with the remaining backward kernels."""
        import torch.distributed as dist

        sv = self._saved
        if sv is None:
            raise RuntimeError("backward() needs a preceding forward_train()")
        overlap = overlap_reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if overlap and accumulate:
            raise ValueError("overlap_reduce reduces fresh gradients: use it on the last micro-batch only, without accumulate")
        self._reduced = False
        B, T, d = sv["B"], sv["T"], self.d
        M = B * T
        grad_out = grad_out.to(device=self.device, dtype=torch.float32).contiguous()
        assert tuple(grad_out.shape) == (T, B, d)
        gtext = self.buf("b_gtext", (M, d), torch.float32)
        if not accumulate and not sv["fused"]:
            # a batch without image features computes no fusion gradients: they must read 0, not the previous batch's
            self.flat_g[self.bucket_top[0]:].zero_()
        if sv["fused"]:
            with _scope("fusion"):
                n_types = len(sv["imgs"])
                if not accumulate and n_types < len(self.fusion):
                    # image types absent from this batch compute no gradients: theirs must read 0
                    self.flat_g[self.bucket_top[0]:].zero_()
                for j in range(n_types):
                    # every type sees the same d loss / d res (the fused states are their sum); d loss / d text adds up
                    gt = gtext if j == 0 else self.buf("b_gtext_j", (M, d), torch.float32)
                    self._fusion_bwd(j, grad_out, gt, B, T, accumulate, ln_accumulate=j > 0)
                    if j > 0:
                        K.dropout(gt, gtext, 0.0, 0, 0, resid=gtext)       # gtext += gt
                if sv["p_text"] > 0:
                    K.dropout(gtext, gtext, sv["p_text"], sv["drop"][2], SITE_TEXT, seed_dev=sv["drop"][3])
        else:
            K.tbc_to_btc(grad_out, B, T, d, gtext)
        g = self.buf("b_g", (M, d), torch.float32)
        g_op = self.buf("b_g_op", (M, d), self.op_dtype)
        lnp = self._lnp()
        top_drop = (sv["drop"][0], sv["drop"][2], sv["drop"][3], site_layer(self.n_layers - 1, 2)) if sv["drop"][0] > 0 else None
        K.layernorm_bwd(sv["x_final"], self.ln_g, gtext, lnp, dx=g, dx_op=g_op, drop=top_drop)
        self._ln_param_grads(lnp, d, self.g(self.enc.layer_norm.weight, self.enc.layer_norm.bias), accumulate)
        self._flush()           # fusion + final LayerNorm reductions
        flush_every = self.wgrad_flush_layers or (3 if overlap else 0)
        if overlap:
            self._wgrad_flush()
            self._reduce_async(*self.bucket_top)
        pending = []
        for i in reversed(range(self.n_layers)):
            with _scope("layer"):
                g_op = self._layer_bwd(i, g, g_op, B, T, sv["seq_lens"], accumulate)
            pending.append(i)
            if not self.grouped_wgrad or (flush_every and len(pending) >= flush_every) or i == 0:
                with _scope("layer"):
                    self._wgrad_flush()
                    self._flush()       # the pending layers' LayerNorm-parameter partials: one launch per 16 reductions
                if overlap:     # the pending layers' buckets are adjacent in flat_g: one exchange
                    self._reduce_async(min(self.bucket_layers[j][0] for j in pending),
                                       max(self.bucket_layers[j][1] for j in pending))
                pending = []
        with _scope("conv"):
            if sv["drop"][0] > 0:      # dropout after the scaled, position-added subsampler output
                K.dropout(g, g, sv["drop"][0], sv["drop"][2], SITE_EMBED, seed_dev=sv["drop"][3])
            self._conv_bwd(g, B, T, accumulate)
            self._flush()
        if overlap:
            self._reduce_async(*self.bucket_conv)
            self._reduce_join()

    # ------------------------------------------------------------------------------------------
    # gradient exchange + optimizer
    # ------------------------------------------------------------------------------------------
    def all_reduce_grads(self, bucket_elems: int = 0) -> int:
        """Sum ``flat_g`` over the ranks of the default process group (one collective; buckets on request); returns the
        world size (fairseq then multiplies the gradients by world_size / sample_size: pass that as ``grad_scale`` to
        ``adam_step``).  The only collective of the path: NCCL over NVLink on the GPU box, gloo in
        the CPU tests (``all_reduce_flat``).  A no-op when ``backward(overlap_reduce=True)`` already reduced them."""
        import torch.distributed as dist

        if self._reduced:
            self._reduced = False
            return dist.get_world_size() if dist.is_initialized() else 1
        return all_reduce_flat(self.flat_g, bucket_elems)

    def adam_step(self, lr: float, betas=(0.9, 0.98), eps: float = 1e-8, weight_decay: float = 0.0,
                  clip_norm: float = 0.0, grad_scale: float = 1.0, extra_norm: Optional[torch.Tensor] = None) -> None:
        """fairseq: multiply_grads(grad_scale) -> clip_grad_norm_(clip_norm) -> Adam.step, then refresh operand copies.
        extra_norm[0]: scaled gradient norm of the rest of the model (the decoder engine's ``grad_norm()``), so that
        the clip coefficient in ``norm_coef[1]`` is the whole model's, as in fairseq."""
        self.step_count += 1
        with _scope("optim"):
            K.grad_clip_coef(self.flat_g, grad_scale, clip_norm, self._sumsq_partials, self.norm_coef,
                             extra_norm=extra_norm)
            K.adam(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, lr=lr, betas=betas, eps=eps,
                   weight_decay=weight_decay, step=self.step_count, norm_coef=self.norm_coef, param_op=self.flat_op)
            self.repack()

    def hyper_values(self, lr: float, betas, weight_decay: float, clip_norm: float, grad_scale: float) -> List[float]:
        """The six per-step floats of ``norm_coef[2:6]`` for the NEXT optimizer step (advances the step counter)."""
        import numpy as np

        self.step_count += 1
        t = self.step_count
        # the same arithmetic as mm_adam's host side (fp32 arguments, double bias corrections): replay == eager, bit for bit
        b1, b2, lr32 = float(np.float32(betas[0])), float(np.float32(betas[1])), float(np.float32(lr))
        step_size = lr32 * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
        return [step_size, float(np.float32(weight_decay) * np.float32(lr)), grad_scale, clip_norm]

    def adam_step_device_hyper(self, betas=(0.9, 0.98), eps: float = 1e-8, extra_norm: Optional[torch.Tensor] = None) -> None:
        """Same optimizer step with step_size / wd*lr / grad_scale / clip_norm read from ``norm_coef[2:6]`` on the device:
        the form that is captured into a CUDA graph (``graph.GraphedTrainStep`` writes the values before each replay)."""
        with _scope("optim"):
            K.grad_clip_coef(self.flat_g, 1.0, 0.0, self._sumsq_partials, self.norm_coef, dev_hyper=True,
                             extra_norm=extra_norm)
            K.adam(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, lr=0.0, betas=betas, eps=eps,
                   weight_decay=0.0, step=0, norm_coef=self.norm_coef, param_op=self.flat_op)
            self.repack()
