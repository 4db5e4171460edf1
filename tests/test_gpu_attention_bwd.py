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


@pytest.mark.parametrize("B,H,T,ragged", [(2, 4, 250, True), (3, 8, 256, False), (2, 2, 100, True), (1, 4, 128, False),
                                           (2, 4, 130, True), (64, 8, 250, True)])
def test_fused_attention_backward_on_chip(cuda, B, H, T, ragged):
    """mm_attention_bwd_fused (S, dP, P, dS on chip; dq | dk | dv accumulated in TMEM) against autograd over the same
    16-bit q | k | v in fp32, and against the two-kernel path it replaces (mm_attention_bwd_scores + mm_heads_gemm)."""
    from mm_s2ut_b200 import kernels as K

    d = 64 * H
    g = torch.Generator().manual_seed(T * 7 + B)
    qkv = (torch.randn(B * T, 3 * d, generator=g) * 0.5).bfloat16().cuda()
    qkv[:, :d] *= 0.125                                           # q arrives pre-scaled by head_dim^-0.5
    dO = (torch.randn(B * T, d, generator=g) * 0.1).bfloat16().cuda()
    lens = torch.tensor([T - (7 * b) % (T // 2) if ragged else T for b in range(B)], dtype=torch.int32, device=cuda)
    out = torch.empty(B * T, d, dtype=torch.bfloat16, device=cuda)
    lse = torch.empty(B, H, T, device=cuda)
    K.self_attention(qkv, lens, B, T, H, out, lse=lse)
    dqkv = torch.full((B * T, 3 * d), float("nan"), dtype=torch.bfloat16, device=cuda)
    K.attention_bwd_fused(qkv, T, lens, B, H, dO, out, lse, dqkv)
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all()
    # ---- the two-kernel path
    Tp = (T + 63) // 64 * 64
    P = torch.zeros(B * H, Tp, Tp, dtype=torch.bfloat16, device=cuda)
    dS = torch.zeros_like(P)
    ref2 = torch.zeros_like(dqkv)
    K.attention_bwd_scores(qkv, 0, T, qkv, d, qkv, 2 * d, T, lens, B, H, dO, out, lse, P, dS)
    hg = dict(a_ld=Tp, a_bs=Tp * Tp, out_ld=3 * d, out_bs=T * 3 * d, rows=T, k=T, batch=B, heads=H)
    K.heads_gemm(P, transposed=True, w=dO, w_ld=d, w_bs=T * d, out=ref2[:, 2 * d:], **hg)
    K.heads_gemm(dS, transposed=True, w=qkv, w_ld=3 * d, w_bs=T * 3 * d, out=ref2[:, d:], **hg)
    K.heads_gemm(dS, transposed=False, w=qkv[:, d:], w_ld=3 * d, w_bs=T * 3 * d, out=ref2, scale=0.125, **hg)
    err2 = (dqkv.float() - ref2.float()).abs().max().item()
    assert err2 <= 2e-2 * max(1.0, ref2.float().abs().max().item()), err2
    if B > 8:
        return
    # ---- autograd in fp32 over the same operands (q un-scaled for autograd, scale applied inside)
    x = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4).clone()           # [3, B, H, T, 64]
    q, k, v = (x[i].detach().requires_grad_(True) for i in range(3))
    s = q @ k.transpose(-1, -2)                                                     # q already scaled
    mask = torch.arange(T, device=cuda)[None, :] >= lens[:, None]
    p = s.masked_fill(mask[:, None, None, :], float("-inf")).softmax(-1)
    o = p @ v
    do = dO.float().view(B, T, H, 64).permute(0, 2, 1, 3)
    (o * do).sum().backward()
    back = lambda t: t.permute(0, 2, 1, 3).reshape(B * T, d)
    ref = torch.cat([back(q.grad), back(k.grad), back(v.grad)], 1)                 # d/d(scaled q) == dS k; ours is * 1/8
    ref[:, :d] *= 0.125
    for name, c0 in (("dq", 0), ("dk", d), ("dv", 2 * d)):
        got, want = dqkv[:, c0:c0 + d].float(), ref[:, c0:c0 + d]
        rel = ((got - want).norm() / want.norm().clamp_min(1e-9)).item()
        assert rel < 2e-2, (name, rel)


@pytest.mark.parametrize("B,H,Lq,Tk,causal,ragged", [(2, 4, 500, 500, True, False), (2, 4, 500, 250, False, True),
                                                      (3, 2, 130, 130, True, False), (2, 4, 750, 750, False, True),
                                                      (1, 8, 300, 64, False, False), (2, 2, 257, 257, True, False),
                                                      (16, 8, 500, 500, True, False)])
def test_general_attention_backward_on_chip(cuda, B, H, Lq, Tk, causal, ragged):
    """mm_attention_bwd_general (query-tile pairs, fp32 dk / dv partials of the earlier pairs in scratch, causal steps
    skipped) against autograd in fp32 over the same 16-bit operands: the decoder's causal self-attention (L = 500), its
    encoder attention (500 x 250 with encoder lengths), the encoder at 30 s (T = 750)."""
    from mm_s2ut_b200 import kernels as K

    d = 64 * H
    g = torch.Generator().manual_seed(Lq * 3 + Tk + B)
    q = (torch.randn(B * Lq, d, generator=g) * 0.5 * 0.125).bfloat16().cuda()         # pre-scaled by head_dim^-0.5
    kv = (torch.randn(B * Tk, 2 * d, generator=g) * 0.5).bfloat16().cuda()
    dO = (torch.randn(B * Lq, d, generator=g) * 0.1).bfloat16().cuda()
    lens = None
    if ragged:
        lens = torch.tensor([Tk - (11 * b) % (Tk // 2) for b in range(B)], dtype=torch.int32, device=cuda)
    out = torch.empty(B * Lq, d, dtype=torch.bfloat16, device=cuda)
    lse = torch.empty(B, H, Lq, device=cuda)
    K.attention(q, 0, Lq, kv, 0, kv, d, Tk, lens, B, H, out, causal=causal, lse=lse)
    dq = torch.full((B * Lq, d), float("nan"), dtype=torch.bfloat16, device=cuda)
    dkv = torch.full((B * Tk, 2 * d), float("nan"), dtype=torch.bfloat16, device=cuda)
    scratch = torch.empty(K.attention_bwd_general_scratch_floats(Tk), device=cuda)
    K.attention_bwd_general(q, Lq, kv, kv[:, d:], Tk, lens, B, H, dO, out, lse, dq, dkv, dkv[:, d:], scratch, causal=causal)
    torch.cuda.synchronize()
    assert torch.isfinite(dq.float()).all() and torch.isfinite(dkv.float()).all()
    if B > 8:
        # twice: the scratch partials of the first run must not leak into the second
        dq2, dkv2 = torch.empty_like(dq), torch.empty_like(dkv)
        K.attention_bwd_general(q, Lq, kv, kv[:, d:], Tk, lens, B, H, dO, out, lse, dq2, dkv2, dkv2[:, d:], scratch, causal=causal)
        assert torch.equal(dq, dq2) and torch.equal(dkv, dkv2)
        return
    heads = lambda x, L: x.float().view(B, L, H, 64).permute(0, 2, 1, 3).clone()
    qf, kf, vf = heads(q, Lq).requires_grad_(), heads(kv[:, :d], Tk).requires_grad_(), heads(kv[:, d:], Tk).requires_grad_()
    s = qf @ kf.transpose(-1, -2)
    if lens is not None:
        s = s.masked_fill((torch.arange(Tk, device=cuda)[None, :] >= lens[:, None])[:, None, None, :], float("-inf"))
    if causal:
        s = s.masked_fill(torch.arange(Tk, device=cuda)[None, :] > torch.arange(Lq, device=cuda)[:, None], float("-inf"))
    o = s.softmax(-1) @ vf
    (o * heads(dO, Lq)).sum().backward()
    back = lambda t, L: t.permute(0, 2, 1, 3).reshape(B * L, d)
    for name, got, want in (("dq", dq, back(qf.grad, Lq) * 0.125), ("dk", dkv[:, :d], back(kf.grad, Tk)),
                            ("dv", dkv[:, d:], back(vf.grad, Tk))):
        rel = ((got.float() - want).norm() / want.norm().clamp_min(1e-9)).item()
        assert rel < 2e-2, (name, rel)
