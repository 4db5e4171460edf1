"""mm_attention: separate q / k / v tensors, different query and key lengths, key lengths, causal mask
(the S2UT decoder's self- and encoder-attention shapes) against fp32 torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, B, Tq, Tk, H, lens, causal):
    hd = 64
    qf = q.float().view(B, Tq, H, hd).permute(0, 2, 1, 3)
    kf = k.float().view(B, Tk, H, hd).permute(0, 2, 1, 3)
    vf = v.float().view(B, Tk, H, hd).permute(0, 2, 1, 3)
    s = qf @ kf.transpose(-1, -2)
    if lens is not None:
        mask = torch.arange(Tk, device=q.device)[None, :] >= lens[:, None]
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    if causal:
        s = s + torch.triu(torch.full((Tq, Tk), float("-inf"), device=q.device), 1)
    return (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3).reshape(B * Tq, H * hd)


@pytest.mark.parametrize("Tq,Tk,lens", [(80, 250, [250, 173, 1]), (300, 125, [125, 77, 30]), (7, 600, [600, 129, 128])])
def test_cross_attention(cuda, Tq, Tk, lens):
    from mm_s2ut_b200 import kernels as K

    dt, H = torch.bfloat16, 4
    B, d = len(lens), 4 * 64
    g = torch.Generator().manual_seed(Tq * 1000 + Tk)
    q = (torch.randn(B * Tq, d, generator=g) * 0.8).to(cuda).to(dt)
    kv = torch.randn(B * Tk, 2 * d, generator=g).to(cuda).to(dt)          # k | v as written by one K|V projection
    sl = torch.tensor(lens, dtype=torch.int32, device=cuda)
    out = torch.zeros(B * Tq, d, dtype=dt, device=cuda)
    K.attention(q, 0, Tq, kv, 0, kv, d, Tk, sl, B, H, out)
    torch.cuda.synchronize()
    ref = _ref(q, kv[:, :d], kv[:, d:], B, Tq, Tk, H, sl, False)
    assert (out.float() - ref).abs().max().item() < 3e-2


@pytest.mark.parametrize("T,lens", [(80, None), (300, None), (257, [257, 100]), (128, None)])
def test_causal_self_attention(cuda, T, lens):
    from mm_s2ut_b200 import kernels as K

    dt, H = torch.float16, 8
    B, d = 2, 8 * 64
    g = torch.Generator().manual_seed(T)
    qkv = torch.randn(B * T, 3 * d, generator=g).to(cuda)
    qkv[:, :d] *= 0.7
    qkv = qkv.to(dt)
    sl = None if lens is None else torch.tensor(lens, dtype=torch.int32, device=cuda)
    out = torch.zeros(B * T, d, dtype=dt, device=cuda)
    K.attention(qkv, 0, T, qkv, d, qkv, 2 * d, T, sl, B, H, out, causal=True)
    torch.cuda.synchronize()
    ref = _ref(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], B, T, T, H, sl, True)
    valid = torch.ones(B, T, dtype=torch.bool, device=cuda)
    if sl is not None:   # rows past an utterance's length attend to nothing meaningful: compare valid rows only
        valid = torch.arange(T, device=cuda)[None, :] < sl[:, None]
    err = (out.float() - ref).abs().view(B, T, d)[valid].max().item()
    assert err < 1e-2, err
