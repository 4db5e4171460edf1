"""mm_attention_bwd_scores (+ the log-sum-exp output of the forward attention kernels) against fp32 torch autograd:
P = softmax(q k^T + masks) and dS = dL/dS for L = <dO, P v>, at the encoder's self-attention shape (T = 250, key
lengths), the decoder's causal self-attention and its encoder attention (query length != key length)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(q, k, v, dO, B, Lq, Tk, H, lens, causal):
    hd = 64
    qf = q.float().view(B, Lq, H, hd).permute(0, 2, 1, 3)
    kf = k.float().view(B, Tk, H, hd).permute(0, 2, 1, 3)
    vf = v.float().view(B, Tk, H, hd).permute(0, 2, 1, 3)
    s = (qf @ kf.transpose(-1, -2)).requires_grad_()
    dead = torch.zeros(B, 1, Lq, Tk, dtype=torch.bool, device=q.device)
    if lens is not None:
        dead = dead | (torch.arange(Tk, device=q.device)[None, :] >= lens[:, None])[:, None, None, :]
    if causal:
        dead = dead | (torch.arange(Tk, device=q.device)[None, :] > torch.arange(Lq, device=q.device)[:, None])[None, None]
    p = torch.softmax(s.masked_fill(dead, float("-inf")), -1)
    o = p @ vf
    dOf = dO.float().view(B, Lq, H, hd).permute(0, 2, 1, 3)
    (o * dOf).sum().backward()
    lse = torch.logsumexp(s.detach().masked_fill(dead, float("-inf")), -1)
    return p.detach(), s.grad, o.detach().permute(0, 2, 1, 3).reshape(B * Lq, H * hd), lse


@pytest.mark.parametrize("B,Lq,Tk,H,lens,causal,dt", [
    (3, 250, 250, 8, [250, 173, 64], False, torch.bfloat16),       # encoder self-attention at the bench length
    (2, 300, 300, 4, None, True, torch.bfloat16),                   # decoder causal self-attention (3 key chunks)
    (2, 300, 125, 4, [125, 77], False, torch.float16),              # decoder -> encoder attention
    (40, 50, 50, 8, None, False, torch.bfloat16),                   # more items than SMs, a single partial chunk
    (1, 130, 130, 4, [3], True, torch.bfloat16),                    # almost everything masked: rows of 1-3 keys
])
def test_scores_backward_matches_autograd(cuda, B, Lq, Tk, H, lens, causal, dt):
    from mm_s2ut_b200 import kernels as K

    d = H * 64
    g = torch.Generator().manual_seed(B * 1000 + Lq + Tk)
    self_attn = Lq == Tk
    if self_attn:      # q | k | v as column blocks of one tensor, like the QKV projection writes them
        qkv = torch.randn(B * Lq, 3 * d, generator=g).to(cuda)
        qkv[:, :d] *= 0.35
        qkv = qkv.to(dt)
        q, k, v, qc, kc, vc = qkv, qkv, qkv, 0, d, 2 * d
        q_, k_, v_ = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    else:
        q = (torch.randn(B * Lq, d, generator=g) * 0.35).to(cuda).to(dt)
        kv = torch.randn(B * Tk, 2 * d, generator=g).to(cuda).to(dt)
        k, v, qc, kc, vc = kv, kv, 0, 0, d
        q_, k_, v_ = q, kv[:, :d], kv[:, d:]
    dO = torch.randn(B * Lq, d, generator=g).to(cuda).to(dt)
    sl = None if lens is None else torch.tensor(lens, dtype=torch.int32, device=cuda)
    # forward kernel: output + log-sum-exp
    out = torch.zeros(B * Lq, d, dtype=dt, device=cuda)
    lse = torch.full((B, H, Lq), float("nan"), dtype=torch.float32, device=cuda)
    if self_attn and not causal:
        K.self_attention(qkv, sl if sl is not None else torch.full((B,), Lq, dtype=torch.int32, device=cuda), B, Lq, H,
                         out, lse=lse)
    else:
        K.attention(q, qc, Lq, k, kc, v, vc, Tk, sl, B, H, out, causal=causal, lse=lse)
    p_ref, ds_ref, o_ref, lse_ref = _reference(q_, k_, v_, dO, B, Lq, Tk, H, sl, causal)
    torch.cuda.synchronize()
    assert (lse - lse_ref).abs().max().item() < 2e-3
    assert (out.float() - o_ref).abs().max().item() < 3e-2
    Lp, Tp = (Lq + 63) // 64 * 64, (Tk + 63) // 64 * 64
    P = torch.full((B * H, Lp, Tp), float("nan"), dtype=dt, device=cuda)
    dS = torch.full((B * H, Lp, Tp), float("nan"), dtype=dt, device=cuda)
    K.attention_bwd_scores(q, qc, Lq, k, kc, v, vc, Tk, sl, B, H, dO, out, lse, P, dS, causal=causal)
    torch.cuda.synchronize()
    Pg = P.view(B, H, Lp, Tp)[:, :, :Lq, :Tk].float()
    dSg = dS.view(B, H, Lp, Tp)[:, :, :Lq, :Tk].float()
    assert torch.isfinite(Pg).all() and torch.isfinite(dSg).all()
    assert (Pg - p_ref).abs().max().item() < 6e-3
    # dS: 16-bit output, delta from the 16-bit forward output; compare in relative L2 over the whole tensor + max-abs
    rel = ((dSg - ds_ref).norm() / ds_ref.norm().clamp_min(1e-6)).item()
    assert rel < 2e-2, rel
    # (rows with one visible key have dS = dP - delta = 0 exactly; delta from the 16-bit O leaves ~|dO| |v| 2^-9 there)
    assert (dSg - ds_ref).abs().max().item() < 6e-2 * max(1.0, ds_ref.abs().max().item())
    dead = p_ref == 0
    assert (Pg[dead] == 0).all() and (dSg[dead] == 0).all()      # masked cells are exact zeros
