"""Operand layouts the backward pass adds to the tcgen05 GEMM (``mm_gemm``): MN-major A / W operands (dgrad and wgrad
read activations, gradients and weights as stored, no transposed copies), split-K batches over the contraction index,
and (sequence, head) batch decomposition straight from / into the q|k|v layout.  Reference: fp32 matmul of the same
16-bit operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).bfloat16().cuda()


def _close(got, ref, tol=2e-2):
    err = (got.float() - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("M,N,Kc", [(300, 320, 192), (1000, 64, 512), (77, 1536, 64)])
def test_dgrad_w_mn_major(cuda, M, N, Kc):
    """out[m, n] = sum_c A[m, c] W[c, n]: W stored [contraction, n] (PyTorch Linear weight [out=c, in=n] as is)."""
    from mm_s2ut_b200 import kernels as K

    A, W = _rnd(M, Kc, seed=1), _rnd(Kc, N, seed=2)
    out = torch.zeros(M, N, dtype=torch.float32, device=cuda)
    K.gemm(a0=A, a0_ld=Kc, rows=M, w=W, w_ld=N, n=N, k=Kc, mode=K.EPI_F32, out0=out, out0_ld=N, w_mn=True)
    _close(out, A.float() @ W.float())
    out_op = torch.zeros(M, N, dtype=torch.bfloat16, device=cuda)
    K.gemm(a0=A, a0_ld=Kc, rows=M, w=W, w_ld=N, n=N, k=Kc, mode=K.EPI_OP, out0=out_op, out0_ld=N, w_mn=True)
    _close(out_op, A.float() @ W.float(), 3e-2)


@pytest.mark.parametrize("M,N,Kin,S", [(1000, 384, 200, 4), (4096, 512, 512, 8), (130, 64, 1024, 1)])
def test_wgrad_both_mn_major_split_k(cuda, M, N, Kin, S):
    """dW[n, k] = sum_m dY[m, n] X[m, k], contraction over tokens split into S batches of `chunk` tokens."""
    from mm_s2ut_b200 import kernels as K

    dY, X = _rnd(M, N, seed=3), _rnd(M, Kin, seed=4)
    chunk = ((M + S - 1) // S + 63) // 64 * 64
    part = torch.full((S, N, Kin), 7.0, dtype=torch.float32, device=cuda)
    K.gemm(a0=dY, a0_ld=N, rows=N, batches=S, w=X, w_ld=Kin, w_batched=True, n=Kin, k=chunk, mode=K.EPI_F32, out0=part,
           out0_ld=Kin, out0_bs=N * Kin, a_mn=True, w_mn=True, a_kbatch=True, w_kbatch=True, a_k_total=M, w_k_total=M)
    _close(part.sum(0), dY.float().t() @ X.float(), 2e-2)


def test_attention_backward_gemms_head_mode(cuda):
    from mm_s2ut_b200 import kernels as K

    B, H, T = 3, 4, 50
    d, Tp = 64 * H, 64
    qkv = _rnd(B * T, 3 * d, seed=5)
    dO = _rnd(B * T, d, seed=6)
    q, k, v = (qkv[:, i * d:(i + 1) * d].float().view(B, T, H, 64).permute(0, 2, 1, 3) for i in range(3))   # [B,H,T,64]
    do = dO.float().view(B, T, H, 64).permute(0, 2, 1, 3)
    # S = q k^T and dP = dO v^T per (sequence, head), straight from the q|k|v layout
    S = torch.zeros(B * H, Tp, Tp, device=cuda)
    K.gemm(a0=qkv, a0_ld=3 * d, a0_bs=T * 3 * d, rows=T, batches=B * H, w=qkv[:, d:], w_ld=3 * d, w_bs=T * 3 * d,
           w_batched=True, n=T, k=64, mode=K.EPI_F32, out0=S, out0_ld=Tp, out0_bs=Tp * Tp, a_hm=True, w_hm=True,
           heads=H, head_stride=64)
    _close(S.view(B, H, Tp, Tp)[:, :, :T, :T], q @ k.transpose(-1, -2))
    dP = torch.zeros(B * H, Tp, Tp, device=cuda)
    K.gemm(a0=dO, a0_ld=d, a0_bs=T * d, rows=T, batches=B * H, w=qkv[:, 2 * d:], w_ld=3 * d, w_bs=T * 3 * d,
           w_batched=True, n=T, k=64, mode=K.EPI_F32, out0=dP, out0_ld=Tp, out0_bs=Tp * Tp, a_hm=True, w_hm=True,
           heads=H, head_stride=64)
    _close(dP.view(B, H, Tp, Tp)[:, :, :T, :T], do @ v.transpose(-1, -2))
    # dV = P^T dO, dK = dS^T q (A MN-major from [q][k] buffers), dQ = dS k (W MN-major), into the q|k|v layout
    P = torch.full((B * H, Tp, Tp), float("nan"), dtype=torch.bfloat16, device=cuda)   # padding is never read
    P[:, :T, :T] = _rnd(B * H, T, T, seed=7)
    Pf = P.float().view(B, H, Tp, Tp)[:, :, :T, :T]
    dqkv = torch.zeros(B * T, 3 * d, dtype=torch.bfloat16, device=cuda)
    og = dict(rows=T, batches=B * H, w_batched=True, n=64, mode=K.EPI_OP, out0_ld=3 * d, out0_bs=T * 3 * d, out_hm=True,
              heads=H, head_stride=64, w_mn=True, w_hm=True)
    # k = T: contraction rows T .. 63 of the last k-block are out of bounds for every operand -> zero-filled by TMA
    K.gemm(a0=P, a0_ld=Tp, a0_bs=Tp * Tp, a_mn=True, k=T, w=dO, w_ld=d, w_bs=T * d, out0=dqkv[:, 2 * d:], **og)
    K.gemm(a0=P, a0_ld=Tp, a0_bs=Tp * Tp, a_mn=True, k=T, w=qkv, w_ld=3 * d, w_bs=T * 3 * d, out0=dqkv[:, d:], **og)
    K.gemm(a0=P, a0_ld=Tp, a0_bs=Tp * Tp, k=T, w=qkv[:, d:], w_ld=3 * d, w_bs=T * 3 * d, out0=dqkv, scale=0.125,
           scale_cols=64, **og)
    back = lambda x: x.permute(0, 2, 1, 3).reshape(B * T, d)
    _close(dqkv[:, 2 * d:], back(Pf.transpose(-1, -2) @ do), 3e-2)
    _close(dqkv[:, d:2 * d], back(Pf.transpose(-1, -2) @ q), 3e-2)
    _close(dqkv[:, :d], back(Pf @ k) * 0.125, 3e-2)


@pytest.mark.parametrize("M,N,Kc,scale", [(300, 2048, 256, 1.0), (1000, 320, 512, 1.25), (77, 64, 64, 1.0),
                                          (16000, 2048, 512, 1.0)])
def test_dgrad_relu_mask_epilogue(cuda, M, N, Kc, scale):
    """MM_EPI_MASK_OP: the fc2 dgrad with the backward of fc1's ReLU (+ activation dropout) in its epilogue:
    out = kept > 0 ? (A W) * scale : 0, `kept` = the 16-bit activation the forward pass saved (autograd of fairseq's
    TransformerEncoderLayer: relu -> activation_dropout -> fc2)."""
    from mm_s2ut_b200 import kernels as K

    A, W = _rnd(M, Kc, seed=11), _rnd(Kc, N, seed=12)
    kept = (_rnd(M, N, seed=13).float().relu() * (torch.rand(M, N, device=cuda) > 0.1)).bfloat16()
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=cuda)
    K.gemm(a0=A, a0_ld=Kc, rows=M, w=W, w_ld=N, n=N, k=Kc, mode=K.EPI_MASK_OP, out0=out, out0_ld=N, w_mn=True,
           aux0=kept, aux_ld=N, scale=scale)
    ref = torch.where(kept.float() > 0, (A.float() @ W.float()) * scale, torch.zeros((), device=cuda))
    _close(out, ref, 3e-2)
    assert torch.equal(out == 0, ~(kept.float() > 0) | (ref.bfloat16() == 0))     # the mask is exact
    assert torch.equal(kept, kept.clone())


@pytest.mark.parametrize("M,N", [(16000, 512), (16000, 2048), (1000, 1536), (130, 64), (577 * 3, 1024)])
def test_colsum_and_batched_reduction(cuda, M, N):
    """Bias gradients: per-chunk column sums of a 16-bit [M, N] matrix + the deterministic partial reduction."""
    from mm_s2ut_b200 import kernels as K

    x = _rnd(M, N, seed=21)
    nb = K.colsum_blocks(M)
    part = torch.full((nb * N,), float("nan"), device=cuda)
    assert K.colsum(x, N, M, N, part) == nb
    out = torch.full((N,), 3.0, device=cuda)
    acc = torch.full((N,), 3.0, device=cuda)
    K.reduce_partials_many([(part, nb, N, N, out, False), (part, nb, N, N, acc, True)])
    ref = x.float().sum(0)
    tol = 1e-3 * max(1.0, ref.abs().max().item())
    assert (out - ref).abs().max().item() <= tol
    assert (acc - 3.0 - ref).abs().max().item() <= tol
    # rows with (r % period) >= valid are skipped
    period, valid = 577, 500
    K.colsum(x, N, M, N, part, period, valid)
    K.reduce_partials_many([(part, nb, N, N, out, False)])
    keep = (torch.arange(M, device=cuda) % period) < valid
    assert (out - (x.float() * keep[:, None]).sum(0)).abs().max().item() <= tol


@pytest.mark.parametrize("accumulate", [False, True])
@pytest.mark.parametrize("tokens", [1000, 16000, 130])
def test_grouped_weight_gradients(cuda, tokens, accumulate):
    """mm_wgrad_grouped: out_g (+)= dy_g^T x_g for several Linear layers in one launch, full token contraction per
    256 x 256 tile (autograd's weight gradient of nn.Linear).  Shapes of one encoder layer + a gate-style column block
    + ragged extents."""
    from mm_s2ut_b200 import kernels as K

    shapes = [(1536, 512), (512, 512), (2048, 512), (512, 2048), (304, 72), (512, 768), (1000, 0)]
    gen = torch.Generator().manual_seed(5)
    groups, refs, outs = [], [], []
    for gi, (n_out, k_in) in enumerate(shapes):
        ld_dy = n_out + (8 if gi == 4 else 0)                 # a strided gradient (column block of a wider tensor)
        dy = (torch.randn(tokens, ld_dy, generator=gen) * 0.5).bfloat16().cuda()
        bias = torch.full((n_out,), 0.25, device=cuda) if gi != 1 else None
        if k_in == 0:                                          # a bias-only group
            groups.append((dy, ld_dy, None, 0, None, 0, n_out, 0, bias))
            refs.append(None)
            outs.append((None, 0, 0, bias, dy[:, :n_out].float().sum(0)))
            continue
        x = torch.randn(tokens, k_in, generator=gen).bfloat16().cuda()
        wide = 2 * k_in if gi == 1 else k_in                   # group 1 fills the second column block of a [n, 2k] gradient
        out = torch.full((n_out, wide), 0.25, device=cuda)
        col = k_in if gi == 1 else 0
        groups.append((dy, ld_dy, x, k_in, out.view(-1)[col:], wide, n_out, k_in, bias))
        refs.append(dy[:, :n_out].float().t() @ x.float())
        outs.append((out, col, k_in, bias, dy[:, :n_out].float().sum(0)))
    K.wgrad_grouped(groups, tokens, accumulate)
    torch.cuda.synchronize()
    base = 0.25 if accumulate else 0.0
    for (out, col, k_in, bias, bref), ref in zip(outs, refs):
        if out is not None:
            got = out[:, col:col + k_in] - base
            err = (got - ref).abs().max().item()
            assert err <= 2e-3 * max(1.0, ref.abs().max().item()) + (1e-3 if accumulate else 0), err
            if col:                                            # the other column block is untouched
                assert torch.equal(out[:, :col], torch.full_like(out[:, :col], 0.25))
        if bias is not None:
            err = (bias - base - bref).abs().max().item()
            assert err <= 2e-3 * max(1.0, bref.abs().max().item()), err


@pytest.mark.parametrize("accumulate", [False, True])
def test_grouped_weight_gradients_of_unequal_token_counts(cuda, accumulate):
    """Groups with their own token counts in ONE launch (the encoder layers' B*T tokens next to the image-side K / V
    projections' B*577): the library orders the tiles by cost (host-side greedy list schedule); more tiles than CTA
    pairs, so the schedule has several rounds and unequal per-pair lists; bias tiles and a bias-only group included."""
    from mm_s2ut_b200 import kernels as K

    #          n_out, k_in, tokens
    shapes = [(1536, 512, 700), (512, 768, 1810), (512, 768, 1810), (2048, 512, 700), (512, 2048, 700), (512, 512, 700),
              (304, 72, 90), (1000, 0, 1810), (1536, 512, 333), (1024, 512, 1234)]
    gen = torch.Generator().manual_seed(9)
    groups, checks = [], []
    for n_out, k_in, tokens in shapes:
        dy = (torch.randn(tokens + 3, n_out, generator=gen) * 0.5).bfloat16().cuda()     # rows beyond `tokens` must not count
        bias = torch.full((n_out,), 0.25, device=cuda)
        bref = dy[:tokens].float().sum(0)
        if k_in == 0:
            groups.append((dy, n_out, None, 0, None, 0, n_out, 0, bias, tokens))
            checks.append((None, None, bias, bref))
            continue
        x = torch.randn(tokens + 3, k_in, generator=gen).bfloat16().cuda()
        out = torch.full((n_out, k_in), 0.25, device=cuda)
        groups.append((dy, n_out, x, k_in, out, k_in, n_out, k_in, bias, tokens))
        checks.append((out, dy[:tokens].float().t() @ x[:tokens].float(), bias, bref))
    assert sum(((n + 255) // 256) * ((k + 255) // 256) for n, k, _ in shapes) > 74
    K.wgrad_grouped(groups, shapes[0][2], accumulate)
    torch.cuda.synchronize()
    base = 0.25 if accumulate else 0.0
    for out, ref, bias, bref in checks:
        if out is not None:
            err = (out - base - ref).abs().max().item()
            assert err <= 2e-3 * max(1.0, ref.abs().max().item()) + (1e-3 if accumulate else 0), err
        err = (bias - base - bref).abs().max().item()
        assert err <= 2e-3 * max(1.0, bref.abs().max().item()), err


@pytest.mark.parametrize("B,H,Lq,Tk", [(3, 4, 50, 50), (2, 8, 250, 250), (2, 2, 500, 250), (1, 4, 130, 577)])
def test_heads_gemm_attention_backward_outputs(cuda, B, H, Lq, Tk):
    """mm_heads_gemm: dV = P^T dO, dK = dS^T q, dQ = dS k * 1/8 per (sequence, head) straight from / into the token-major
    q | k | v layouts, on 128 x 64 tiles (autograd of fairseq MultiheadAttention; same contractions as
    test_attention_backward_gemms_head_mode runs on mm_gemm)."""
    from mm_s2ut_b200 import kernels as K

    d = 64 * H
    Lp, Tp = (Lq + 63) // 64 * 64, (Tk + 63) // 64 * 64
    q = _rnd(B * Lq, d, seed=31)
    kv = _rnd(B * Tk, 2 * d, seed=32)
    dO = _rnd(B * Lq, d, seed=33)
    P = torch.full((B * H, Lp, Tp), float("nan"), dtype=torch.bfloat16, device=cuda)      # padding is never read
    dS = torch.full((B * H, Lp, Tp), float("nan"), dtype=torch.bfloat16, device=cuda)
    P[:, :Lq, :Tk] = _rnd(B * H, Lq, Tk, seed=34) * 0.1
    dS[:, :Lq, :Tk] = _rnd(B * H, Lq, Tk, seed=35) * 0.1
    dq = torch.full((B * Lq, d), float("nan"), dtype=torch.bfloat16, device=cuda)
    dkv = torch.full((B * Tk, 2 * d), float("nan"), dtype=torch.bfloat16, device=cuda)
    hg = dict(a_ld=Tp, a_bs=Lp * Tp, batch=B, heads=H)
    K.heads_gemm(P, transposed=True, w=dO, w_ld=d, w_bs=Lq * d, out=dkv[:, d:], out_ld=2 * d, out_bs=Tk * 2 * d, rows=Tk,
                 k=Lq, **hg)
    K.heads_gemm(dS, transposed=True, w=q, w_ld=d, w_bs=Lq * d, out=dkv, out_ld=2 * d, out_bs=Tk * 2 * d, rows=Tk, k=Lq,
                 **hg)
    K.heads_gemm(dS, transposed=False, w=kv, w_ld=2 * d, w_bs=Tk * 2 * d, out=dq, out_ld=d, out_bs=Lq * d, rows=Lq, k=Tk,
                 scale=0.125, **hg)
    heads = lambda x, L: x.float().view(B, L, H, 64).permute(0, 2, 1, 3)            # [B, H, L, 64]
    back = lambda x, L: x.permute(0, 2, 1, 3).reshape(B * L, d)
    Pf = P[:, :Lq, :Tk].float().view(B, H, Lq, Tk)
    dSf = dS[:, :Lq, :Tk].float().view(B, H, Lq, Tk)
    _close(dkv[:, d:], back(Pf.transpose(-1, -2) @ heads(dO, Lq), Tk), 2e-2)
    _close(dkv[:, :d], back(dSf.transpose(-1, -2) @ heads(q, Lq), Tk), 2e-2)
    _close(dq, back(dSf @ heads(kv[:, :d], Tk), Lq) * 0.125, 2e-2)


@pytest.mark.parametrize("rows,k,p_drop", [(700, 1536, 0.0), (16000, 2048, 0.1), (333, 512, 0.25), (40000, 512, 0.0)])
def test_dgrad_layernorm_backward_fused(cuda, rows, k, p_drop):
    """mm_gemm_ln_bwd: dgrad of a Linear after a LayerNorm + the LayerNorm backward + the residual add in one kernel,
    against the same steps in fp32 (autograd's formulas; the un-fused path is mm_gemm(EPI_F32) + mm_layernorm_bwd_drop)."""
    from mm_s2ut_b200 import kernels as K

    gen = torch.Generator().manual_seed(rows + k)
    dy = (torch.randn(rows, k, generator=gen) * 0.05).bfloat16().cuda()
    w = (torch.randn(k, 512, generator=gen) * 0.05).bfloat16().cuda()
    x = (torch.randn(rows, 512, generator=gen) * 1.7 + 0.4).cuda()
    gamma = (1.0 + 0.2 * torch.randn(512, generator=gen)).cuda()
    g0 = (torch.randn(rows, 512, generator=gen) * 0.1).cuda()
    g = g0.clone()
    g_op = torch.full((rows, 512), float("nan"), dtype=torch.bfloat16, device=cuda)
    nrows = K.gemm_ln_bwd_partial_rows(rows)
    part = torch.full((nrows * 1024,), float("nan"), device=cuda)
    drop = (p_drop, 77, None, 5) if p_drop > 0 else None
    assert K.gemm_ln_bwd(dy, w, x, gamma, g, g_op, part, drop=drop) == nrows
    torch.cuda.synchronize()
    dh = dy.float() @ w.float()
    mean = x.mean(1, keepdim=True)
    rstd = torch.rsqrt(((x - mean) ** 2).mean(1, keepdim=True) + 1e-5)
    xhat = (x - mean) * rstd
    v = dh * gamma
    ref = g0 + rstd * (v - v.mean(1, keepdim=True) - xhat * (v * xhat).mean(1, keepdim=True))
    scale = ref.abs().max().item()
    assert (g - ref).abs().max().item() <= 4e-3 * scale            # (dh, xhat) pass through bf16 between the two sweeps
    rel = ((g - ref).norm() / (ref - g0).norm()).item()
    assert rel < 6e-3, rel
    keep = torch.ones_like(g)
    if p_drop > 0:
        K.dropout(keep, keep, p_drop, 77, 5)
    assert torch.equal(g_op, (g * keep).bfloat16())
    sums = part.view(nrows, 2, 512).sum(0)
    for got, want in ((sums[0], (dh * xhat).sum(0)), (sums[1], dh.sum(0))):
        assert (got - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())
