"""tcgen05 GEMM + fused epilogues through the C ABI vs plain PyTorch fp32 on the same 16-bit operands."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, dev, dt, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).to(dt)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("bn", [256])
@pytest.mark.parametrize("M,N,Kd", [(300, 512, 512), (1000, 768, 256), (128, 256, 2048), (77, 1536, 512)])
def test_gemm_bias_op(cuda, dt, bn, M, N, Kd):
    from mm_s2ut_b200 import kernels as K

    a, w = _rand((M, Kd), cuda, dt, 1.0, 1), _rand((N, Kd), cuda, dt, Kd ** -0.5, 2)
    bias = _rand((N,), cuda, torch.float32, 1.0, 3)
    out = torch.zeros(M, N, dtype=dt, device=cuda)
    K.gemm(a0=a, a0_ld=Kd, rows=M, w=w, n=N, k=Kd, mode=K.EPI_OP, bias=bias, out0=out, out0_ld=N, scale=0.125,
           scale_cols=256, block_n=bn)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    ref[:, :256] *= 0.125
    err = (out.float() - ref).abs().max().item()
    assert err < 3e-2 if dt == torch.bfloat16 else err < 5e-3, err


def test_gemm_relu_resid_f32(cuda):
    from mm_s2ut_b200 import kernels as K

    dt, M, N, Kd = torch.bfloat16, 1000, 512, 2048
    a, w = _rand((M, Kd), cuda, dt, 1.0, 1), _rand((N, Kd), cuda, dt, Kd ** -0.5, 2)
    bias = _rand((N,), cuda, torch.float32, 1.0, 3)
    x = _rand((M, N), cuda, torch.float32, 1.0, 4)
    ref = x + a.float() @ w.float().t() + bias
    K.gemm(a0=a, a0_ld=Kd, rows=M, w=w, n=N, k=Kd, mode=K.EPI_RESID_F32, bias=bias, aux0=x, aux_ld=N, out0=x,
           out0_ld=N)
    torch.cuda.synchronize()
    assert (x - ref).abs().max().item() < 2e-3
    out = torch.zeros(M, N, dtype=dt, device=cuda)
    K.gemm(a0=a, a0_ld=Kd, rows=M, w=w, n=N, k=Kd, mode=K.EPI_RELU_OP, bias=bias, out0=out, out0_ld=N)
    torch.cuda.synchronize()
    ref = torch.relu(a.float() @ w.float().t() + bias)
    assert (out.float() - ref).abs().max().item() < 3e-2


def test_gemm_vt_and_rows_per_seq(cuda):
    """QKV-style epilogue: q scaled, k plain, v stored transposed per sequence."""
    from mm_s2ut_b200 import kernels as K

    dt, B, T, d = torch.bfloat16, 3, 250, 512
    M, Tp = B * T, 256
    a, w = _rand((M, d), cuda, dt, 1.0, 1), _rand((3 * d, d), cuda, dt, d ** -0.5, 2)
    bias = _rand((3 * d,), cuda, torch.float32, 0.1, 3)
    qk = torch.zeros(M, 2 * d, dtype=dt, device=cuda)
    vt = torch.zeros(B, d, Tp, dtype=dt, device=cuda)
    K.gemm(a0=a, a0_ld=d, rows=M, w=w, n=3 * d, k=d, mode=K.EPI_OP, bias=bias, scale=0.125, scale_cols=d, out0=qk,
           out0_ld=2 * d, out0_bs=T * 2 * d, rows_per_seq=T, vt=vt, vt_col0=2 * d, vt_rows=d, vt_ld=Tp)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    ref[:, :d] *= 0.125
    assert (qk.float() - ref[:, : 2 * d]).abs().max().item() < 3e-2
    vref = ref[:, 2 * d:].view(B, T, d).transpose(1, 2)
    assert (vt[:, :, :T].float() - vref).abs().max().item() < 3e-2
    assert vt[:, :, T:].abs().max().item() == 0


def test_gemm_conv_glu_windows(cuda):
    """Stride-2 k=5 Conv1d + GLU as a GEMM over overlapping rows of a time-major buffer (TMA row stride < K)."""
    from mm_s2ut_b200 import kernels as K
    import torch.nn.functional as F

    dt, B, m, cin, cout, bn = torch.bfloat16, 3, 137, 80, 512, 256
    x = _rand((B, m, cin), cuda, dt, 1.0, 1)
    wt = _rand((cout, cin, 5), cuda, dt, (5 * cin) ** -0.5, 2)
    bias = _rand((cout,), cuda, torch.float32, 0.5, 3)
    T1 = (m - 1) // 2 + 1
    m_alloc = m + 4 + ((m + 4) & 1)
    x1 = torch.zeros(B, m_alloc, cin, dtype=dt, device=cuda)
    x1[:, 2: 2 + m] = x
    half = cout // 2
    idx = torch.cat([torch.cat([torch.arange(t * bn // 2, (t + 1) * bn // 2), half + torch.arange(t * bn // 2, (t + 1) * bn // 2)])
                     for t in range(cout // bn)]).to(cuda)
    wg = wt.float().permute(0, 2, 1).reshape(cout, 5 * cin)[idx].to(dt).contiguous()
    T1_alloc = T1 + 4 + ((T1 + 4) & 1)
    out = torch.zeros(B, T1_alloc, half, dtype=dt, device=cuda)
    K.gemm(a0=x1, a0_ld=2 * cin, a0_bs=m_alloc * cin, rows=T1, batches=B, w=wg, n=cout, k=5 * cin, mode=K.EPI_GLU_OP,
           bias=bias[idx].contiguous(), out0=out, out0_ld=half, out0_bs=T1_alloc * half, out_row_offset=2, block_n=bn)
    torch.cuda.synchronize()
    ref = F.glu(F.conv1d(x.float().transpose(1, 2), wt.float(), bias, stride=2, padding=2), dim=1).transpose(1, 2)
    assert ref.shape[1] == T1
    assert (out[:, 2: 2 + T1].float() - ref).abs().max().item() < 3e-2
    assert out[:, :2].abs().max().item() == 0 and out[:, 2 + T1:].abs().max().item() == 0


def test_gemm_glu_pos(cuda):
    from mm_s2ut_b200 import kernels as K
    from mm_s2ut_b200.models.modules import SinusoidalPositionalEmbedding
    import torch.nn.functional as F

    dt, B, T1, cin, d, bn = torch.bfloat16, 2, 61, 512, 256, 256
    cout = 2 * d
    x = _rand((B, T1, cin), cuda, dt, 1.0, 1)
    wt = _rand((cout, cin, 5), cuda, dt, (5 * cin) ** -0.5, 2)
    bias = _rand((cout,), cuda, torch.float32, 0.5, 3)
    T = (T1 - 1) // 2 + 1
    T1_alloc = T1 + 4 + ((T1 + 4) & 1)
    x2 = torch.zeros(B, T1_alloc, cin, dtype=dt, device=cuda)
    x2[:, 2: 2 + T1] = x
    idx = torch.cat([torch.cat([torch.arange(t * bn // 2, (t + 1) * bn // 2), d + torch.arange(t * bn // 2, (t + 1) * bn // 2)])
                     for t in range(cout // bn)]).to(cuda)
    wg = wt.float().permute(0, 2, 1).reshape(cout, 5 * cin)[idx].to(dt).contiguous()
    pos = SinusoidalPositionalEmbedding.get_embedding(T + 2, d, 1).to(cuda)
    lens = torch.tensor([T, T - 7], dtype=torch.int32, device=cuda)
    out = torch.zeros(B * T, d, dtype=torch.float32, device=cuda)
    K.gemm(a0=x2, a0_ld=2 * cin, a0_bs=T1_alloc * cin, rows=T, batches=B, w=wg, n=cout, k=5 * cin,
           mode=K.EPI_GLU_POS_F32, bias=bias[idx].contiguous(), out0=out, out0_ld=d, out0_bs=T * d,
           scale=math.sqrt(d), pos=pos, seq_lens=lens, block_n=bn)
    torch.cuda.synchronize()
    ref = F.glu(F.conv1d(x.float().transpose(1, 2), wt.float(), bias, stride=2, padding=2), dim=1).transpose(1, 2)
    ref = ref * math.sqrt(d)
    p = pos[2: 2 + T].unsqueeze(0).repeat(B, 1, 1)
    p[1, T - 7:] = 0
    ref = ref + p
    assert (out.view(B, T, d) - ref).abs().max().item() < 0.15  # bf16 operands, outputs scaled by sqrt(d)=16


def test_gemm_batched_scores_and_pv(cuda):
    """Batched W (per-utterance keys) with ragged N=577 and K=584 tails: the speech->image attention GEMMs."""
    from mm_s2ut_b200 import kernels as K

    dt, B, T, d, Tk, Tkp = torch.bfloat16, 3, 250, 512, 577, 584
    q, k = _rand((B, T, d), cuda, dt, 1.0, 1), _rand((B, Tk, d), cuda, dt, d ** -0.5, 2)
    S = torch.full((B, T, Tkp), 7.0, dtype=torch.float32, device=cuda)
    K.gemm(a0=q, a0_ld=d, a0_bs=T * d, rows=T, batches=B, w=k, w_ld=d, w_bs=Tk * d, w_batched=True, n=Tk, k=d,
           mode=K.EPI_F32, out0=S, out0_ld=Tkp, out0_bs=T * Tkp)
    torch.cuda.synchronize()
    ref = torch.einsum("btd,bkd->btk", q.float(), k.float())
    assert (S[:, :, :Tk] - ref).abs().max().item() < 2e-3
    # (the TMA store may zero the rest of the last 16-byte group, columns [Tk, Tk+3]; beyond that S is untouched)
    assert (S[:, :, Tk + 3:] == 7.0).all()
    P = torch.zeros(B, T, Tkp, dtype=dt, device=cuda)
    P[:, :, :Tk] = torch.softmax(ref, -1).to(dt)
    vt = torch.zeros(B, d, Tkp, dtype=dt, device=cuda)
    vt[:, :, :Tk] = _rand((B, d, Tk), cuda, dt, 1.0, 5)
    o = torch.zeros(B * T, d, dtype=dt, device=cuda)
    K.gemm(a0=P, a0_ld=Tkp, a0_bs=T * Tkp, rows=T, batches=B, w=vt, w_ld=Tkp, w_bs=d * Tkp, w_batched=True, n=d,
           k=Tkp, mode=K.EPI_OP, out0=o, out0_ld=d, out0_bs=T * d)
    torch.cuda.synchronize()
    oref = torch.einsum("btk,bdk->btd", P.float(), vt.float())
    assert (o.view(B, T, d).float() - oref).abs().max().item() < 1e-2


def test_gemm_f32op_and_gate(cuda):
    from mm_s2ut_b200 import kernels as K

    dt, B, T, d = torch.bfloat16, 3, 125, 256
    M = B * T
    o, wp = _rand((M, d), cuda, dt, 1.0, 1), _rand((d, d), cuda, dt, d ** -0.5, 2)
    bp = _rand((d,), cuda, torch.float32, 0.1, 3)
    a32 = torch.zeros(M, d, dtype=torch.float32, device=cuda)
    a16 = torch.zeros(M, d, dtype=dt, device=cuda)
    K.gemm(a0=o, a0_ld=d, rows=M, w=wp, n=d, k=d, mode=K.EPI_F32_OP, bias=bp, out0=a32, out0_ld=d, out1=a16, out1_ld=d)
    torch.cuda.synchronize()
    ref = o.float() @ wp.float().t() + bp
    assert (a32 - ref).abs().max().item() < 2e-3 and (a16.float() - ref).abs().max().item() < 3e-2
    text32 = _rand((M, d), cuda, torch.float32, 1.0, 4)
    text16 = text32.to(dt)
    wg, bg = _rand((d, 2 * d), cuda, dt, (2 * d) ** -0.5, 5), _rand((d,), cuda, torch.float32, 0.1, 6)
    res = torch.zeros(T, B, d, dtype=torch.float32, device=cuda)
    K.gemm(a0=a16, a0_ld=d, a0_bs=T * d, a1=text16, a1_ld=d, a1_bs=T * d, k_split=d, rows=T, batches=B, w=wg, n=d,
           k=2 * d, mode=K.EPI_GATE, bias=bg, aux0=text32, aux1=a32, aux_ld=d, out0=res, out0_ld=d, out_tbc=True,
           n_seqs=B)
    torch.cuda.synchronize()
    g = torch.sigmoid(torch.cat([a16.float(), text16.float()], -1) @ wg.float().t() + bg)
    rref = ((1 - g) * text32 + g * a32).view(B, T, d).transpose(0, 1)
    assert (res - rref).abs().max().item() < 2e-3
    # residual-sum variant of the fusion output (use_selective_gate: False), also stored T x B x C
    res2 = torch.zeros(T, B, d, dtype=torch.float32, device=cuda)
    K.gemm(a0=o, a0_ld=d, a0_bs=T * d, rows=T, batches=B, w=wp, n=d, k=d, mode=K.EPI_RESID_F32, bias=bp, aux0=text32,
           aux_ld=d, out0=res2, out0_ld=d, out_tbc=True, n_seqs=B)
    torch.cuda.synchronize()
    assert (res2 - (text32 + ref).view(B, T, d).transpose(0, 1)).abs().max().item() < 2e-3


def test_gemm_full_size_linearity(cuda):
    """BASELINE config[1] size (M=16000): size-independent property -- GEMM(a1+a2) == GEMM(a1)+GEMM(a2) up to rounding,
    and spot rows against a torch fp32 reference."""
    from mm_s2ut_b200 import kernels as K

    dt, M, N, Kd = torch.bfloat16, 16000, 2048, 512
    a, w = _rand((M, Kd), cuda, dt, 1.0, 1), _rand((N, Kd), cuda, dt, Kd ** -0.5, 2)
    out = torch.zeros(M, N, dtype=dt, device=cuda)
    K.gemm(a0=a, a0_ld=Kd, rows=M, w=w, n=N, k=Kd, mode=K.EPI_RELU_OP, out0=out, out0_ld=N)
    torch.cuda.synchronize()
    rows = torch.tensor([0, 1, 127, 128, 8191, 15999, 15872], device=cuda)
    ref = torch.relu(a[rows].float() @ w.float().t())
    assert (out[rows].float() - ref).abs().max().item() < 3e-2
    ref_all = torch.relu(a.float() @ w.float().t())
    assert (out.float() - ref_all).abs().max().item() < 3e-2


@pytest.mark.parametrize("M,Kd,want_f32", [(1000, 512, False), (300, 2048, True), (20000, 512, True),
                                           (16000, 512, False), (20000, 2048, False), (40000, 576, False),
                                           (12000, 512, True), (9700, 2048, False)])
def test_gemm_resid_ln_fused(cuda, M, Kd, want_f32):
    """x += a W^T + b and h = LayerNorm(x) in one kernel (20000 rows -> some CTA pairs take two tiles).

    40000 rows -> three row blocks per CTA pair; K=576 exercises nine k-blocks (ring wrap) and 2048 the fc2 shape."""
    from mm_s2ut_b200 import kernels as K

    dt, N = torch.bfloat16, 512
    a, w = _rand((M, Kd), cuda, dt, 1.0, 1), _rand((N, Kd), cuda, dt, Kd ** -0.5, 2)
    bias = _rand((N,), cuda, torch.float32, 0.5, 3)
    x = _rand((M, N), cuda, torch.float32, 2.0, 4) + 0.3
    gamma, beta = 1 + _rand((N,), cuda, torch.float32, 0.2, 5), _rand((N,), cuda, torch.float32, 0.2, 6)
    xref = x + a.float() @ w.float().t() + bias
    href = torch.nn.functional.layer_norm(xref, (N,), gamma, beta, 1e-5)
    h16 = torch.zeros(M, N, dtype=dt, device=cuda)
    h32 = torch.zeros(M, N, dtype=torch.float32, device=cuda) if want_f32 else None
    K.gemm_resid_ln(a, w, bias, x, gamma, beta, h16, h32)
    torch.cuda.synchronize()
    assert (x - xref).abs().max().item() < 2e-3
    assert (h16.float() - href).abs().max().item() < 4e-2
    if want_f32:
        assert (h32 - href).abs().max().item() < 2e-3
