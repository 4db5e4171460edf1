"""mm_cross_attention: the fused (flash-style) speech -> image attention kernel against fp32 torch.

Reference semantics: SelectiveAttention.forward (mm_s2ut/models/fuse.py:80-113, one head of width d_model, optional
key-padding masked_fill(-inf)) and MultimodalAttention (fuse.py:145-167: one extra learned key / value row).  The
shapes are the ones the three BASELINE configurations produce: Tk = 577 / 578 (ViT, +bias_kv) at d = 512 and d = 256,
Tk = 100 (DETR) at d = 1024; query counts that are not multiples of the 128-row tile.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, B, Tq, Tk, mask):
    d = q.shape[1]
    s = q.float().view(B, Tq, d) @ k.float().view(B, Tk, d).transpose(1, 2)
    if mask is not None:
        s = s.masked_fill(mask[:, None, :Tk].bool(), float("-inf"))
    lse = torch.logsumexp(s, -1)
    return (torch.softmax(s, -1) @ v.float().view(B, Tk, d)).reshape(B * Tq, d), lse.reshape(-1)


@pytest.mark.parametrize("B,Tq,Tk,d,dt", [
    (3, 250, 577, 512, torch.bfloat16),      # base model, ViT features (the bench shape per utterance)
    (2, 250, 578, 512, torch.float16),       # MultimodalAttention: + the learned bias_k / bias_v row
    (5, 125, 577, 256, torch.bfloat16),      # small model: a single 256-column block
    (2, 251, 100, 1024, torch.bfloat16),     # large model, DETR features: one key chunk, four column blocks
    (70, 33, 130, 512, torch.bfloat16),      # more items than SMs x 2: the persistent loop and the S / O hand-over
    (1, 1, 1, 256, torch.bfloat16),
])
def test_matches_fp32_softmax_attention(cuda, B, Tq, Tk, d, dt):
    from mm_s2ut_b200 import kernels as K

    g = torch.Generator().manual_seed(B * 7919 + Tq * 31 + Tk)
    q = (torch.randn(B * Tq, d, generator=g) * d ** -0.25).to(cuda).to(dt)        # scores ~ N(0, 1) .. a few units
    kv = torch.randn(B * Tk, 2 * d, generator=g).to(cuda)
    kv[:, :d] *= d ** -0.25
    kv = kv.to(dt)
    out = torch.full((B * Tq, d), float("nan"), dtype=dt, device=cuda)
    lse = torch.full((B * Tq,), float("nan"), dtype=torch.float32, device=cuda)
    K.cross_attention(q, Tq, kv, 0, kv, d, Tk, B, d, out, lse=lse)
    torch.cuda.synchronize()
    ref, lse_ref = _ref(q, kv[:, :d], kv[:, d:], B, Tq, Tk, None)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, err
    assert (lse - lse_ref).abs().max().item() < 2e-3


def test_large_scores_lazy_rescale(cuda):
    """Row maxima that grow chunk after chunk by far more than 2^8: the lazily moved reference maximum must rescale O."""
    from mm_s2ut_b200 import kernels as K

    B, Tq, Tk, d, dt = 2, 130, 577, 512, torch.bfloat16
    g = torch.Generator().manual_seed(11)
    q = torch.randn(B * Tq, d, generator=g).to(cuda).to(dt) * 0.3
    k = torch.randn(B, Tk, d, generator=g).to(cuda) * 0.3
    k *= (1.0 + torch.arange(Tk, device=cuda).float() / 64.0)[None, :, None]    # later keys score ~10x larger
    kv = torch.cat([k.view(B * Tk, d), torch.randn(B * Tk, d, generator=g).to(cuda)], 1).to(dt)
    out = torch.zeros(B * Tq, d, dtype=dt, device=cuda)
    K.cross_attention(q, Tq, kv, 0, kv, d, Tk, B, d, out)
    torch.cuda.synchronize()
    ref, _ = _ref(q, kv[:, :d], kv[:, d:], B, Tq, Tk, None)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() < 3e-2


def test_key_mask(cuda):
    """Image key-padding mask (fuse.py:88-91): masked keys get probability 0, whatever chunk they sit in."""
    from mm_s2ut_b200 import kernels as K

    B, Tq, Tk, d, dt = 3, 140, 300, 512, torch.bfloat16
    g = torch.Generator().manual_seed(5)
    q = (torch.randn(B * Tq, d, generator=g) * d ** -0.25).to(cuda).to(dt)
    kv = (torch.randn(B * Tk, 2 * d, generator=g)).to(cuda)
    kv[:, :d] *= d ** -0.25
    kv = kv.to(dt)
    mask = torch.zeros(B, Tk + 3, dtype=torch.uint8, device=cuda)       # a row stride that is not Tk
    mask[0, 200:] = 1               # a masked tail that covers whole chunks
    mask[1, ::3] = 1                # scattered
    mask[2, :129] = 1               # the whole first chunk masked: the running maximum starts at -inf
    out = torch.zeros(B * Tq, d, dtype=dt, device=cuda)
    K.cross_attention(q, Tq, kv, 0, kv, d, Tk, B, d, out, key_mask=mask)
    torch.cuda.synchronize()
    ref, _ = _ref(q, kv[:, :d], kv[:, d:], B, Tq, Tk, mask)
    assert (out.float() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("attn_type", ["selective_attention", "multimodal_attention"])
def test_encoder_fused_equals_unfused_path(cuda, attn_type):
    """The module forward with the fused kernel against the same forward through scores GEMM -> softmax -> P V GEMM."""
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder

    cfg = dict(load_mm_config(DEFAULT_YAML))
    cfg["multimodal_attention_type"] = attn_type
    torch.manual_seed(3)
    args = make_args("small", multimodal_translation_config_yaml=cfg)
    enc = MM_S2STransformerEncoder(args, build_unused_projections=False).eval().to(cuda)
    wavs, _ = synth.synth_batch(3, 3, 2.0, ragged=True)
    wav, lens = synth.pad_waveforms(wavs)
    imgs = synth.synth_images(3, 3).to(cuda)
    mask = torch.zeros(3, imgs.shape[1], dtype=torch.bool, device=cuda)
    mask[1, 400:] = True
    outs = []
    for fused in (True, False):
        enc.fuse_cross_attention = fused
        enc._engine = None
        for m in (None, mask):
            o = enc(wav.to(cuda), lens.to(cuda), None, None, None, imgs_list=[imgs], img_masks_list=[m])
            outs.append(o["encoder_out"][0].clone())
    torch.cuda.synchronize()
    assert enc.engine().fused_xattn is False
    for a, b in zip(outs[:2], outs[2:]):
        assert (a - b).abs().max().item() < 1e-2
