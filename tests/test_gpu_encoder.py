"""End-to-end parity of the CUDA path (through the module API / C ABI) against the fp32 CPU oracle.

Tolerances are the north_star ones: fused / encoder states within 2e-2 max-abs over valid positions
(bf16 operands, fp32 oracle).
"""
import pytest
import torch

from _util import record

pytestmark = pytest.mark.gpu

TOL = 2e-2


def _build(preset, attn_type="selective_attention", gate=True, seed=0, op_dtype=torch.bfloat16):
    from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder

    cfg = dict(load_mm_config(DEFAULT_YAML))
    cfg["multimodal_attention_type"] = attn_type
    cfg["use_selective_gate"] = gate
    torch.manual_seed(seed)
    args = make_args(preset, multimodal_translation_config_yaml=cfg)
    enc = MM_S2STransformerEncoder(args, build_unused_projections=False).eval()
    if attn_type == "multimodal_attention":  # torch initialises bias_k/v ~ xavier_normal; make them non-trivial
        for m in enc.multimodal_attns:
            torch.nn.init.normal_(m.bias_k, std=0.5)
            torch.nn.init.normal_(m.bias_v, std=0.5)
    # non-trivial LayerNorm affine + biases so that every epilogue term is exercised
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in enc.named_parameters():
            if "layer_norm" in n or "pre_norm" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif n.endswith(".bias") or n.endswith("in_proj_bias"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    enc.op_dtype = op_dtype
    return enc, args, load_mm_config(cfg)


def _oracle(enc, args, cfg, wavs, imgs, img_mask=None, **kw):
    from oracle import fbank as ofb, fusion as ofu

    sd = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    feats, flens = ofb.features_from_waveforms(wavs)
    return ofu.mm_encoder_forward(sd, cfg, torch.from_numpy(feats), torch.from_numpy(flens), [imgs], [img_mask],
                                  args.encoder_attention_heads, **kw)


def _compare(out, ref):
    x, r = out["encoder_out"][0].float().cpu(), ref["encoder_out"][0]
    assert x.shape == r.shape
    mask = ref["encoder_padding_mask"][0] if ref["encoder_padding_mask"] else torch.zeros(r.shape[1], r.shape[0], dtype=torch.bool)
    assert torch.equal(out["encoder_padding_mask"][0].cpu(), mask)
    valid = (~mask).t().unsqueeze(-1)
    err = ((x - r).abs() * valid).max().item()
    assert torch.isfinite(x).all()
    return err


@pytest.mark.parametrize("attn_type,gate", [("selective_attention", True), ("multimodal_attention", True),
                                            ("selective_attention", False)])
def test_small_config0_parity(cuda, attn_type, gate):
    """BASELINE configs[0]: small (6 layers, d=256), B=4 x 5 s ragged + one all-zero utterance."""
    from mm_s2ut_b200 import synth

    enc, args, cfg = _build("small", attn_type, gate)
    wavs, _ = synth.synth_batch(0, 4, 5.0, ragged=True, zero_utt=3)
    imgs = synth.synth_images(0, 4)
    ref = _oracle(enc, args, cfg, wavs, imgs, return_all_hiddens=True)
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda()
    out = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[imgs.cuda()], img_masks_list=[None],
              return_all_hiddens=True)
    torch.cuda.synchronize()
    err = _compare(out, ref)
    record(f"configs[0] small B=4x5s ragged+zero utt, {attn_type}, gate={gate}: fused states max-abs err", err, TOL)
    assert err < TOL, err
    # second forward on the cached workspaces (the learned bias_k / bias_v key is written only when they are created)
    out2 = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[imgs.cuda()], img_masks_list=[None])
    torch.cuda.synchronize()
    assert torch.equal(out2["encoder_out"][0], out["encoder_out"][0])
    # per-layer states (fp32 residual stream): tighter than the final tolerance early on
    mask = ref["encoder_padding_mask"][0]
    valid = (~mask).t().unsqueeze(-1)
    for i, (a, b) in enumerate(zip(out["encoder_states"], ref["encoder_states"])):
        e = ((a.float().cpu() - b).abs() * valid).max().item()
        scale = b.abs().max().item()
        assert e < 5e-3 * max(scale, 1.0) * (i + 2), (i, e, scale)


def test_small_features_input_and_image_mask(cuda):
    """Reference input contract: src_tokens = precomputed [B, T, 80] features; plus an image key-padding mask."""
    from mm_s2ut_b200 import synth
    from oracle import fbank as ofb

    enc, args, cfg = _build("small")
    wavs, _ = synth.synth_batch(0, 3, 4.0, ragged=True)
    imgs = synth.synth_images(0, 3, 100, 768)
    img_mask = torch.zeros(3, 100, dtype=torch.bool)
    img_mask[1, 60:] = True
    ref = _oracle(enc, args, cfg, wavs, imgs, img_mask)
    feats, flens = ofb.features_from_waveforms(wavs)
    enc.cuda()
    out = enc(torch.from_numpy(feats).cuda(), torch.from_numpy(flens).cuda(), None, None, None,
              imgs_list=[imgs.cuda()], img_masks_list=[img_mask.cuda()])
    torch.cuda.synchronize()
    assert _compare(out, ref) < TOL


def test_no_fusion_and_no_padding(cuda):
    """No images -> plain S2T encoder output; equal-length batch -> all-False mask (reference raises IndexError)."""
    from mm_s2ut_b200 import synth

    enc, args, cfg = _build("small")
    wavs, _ = synth.synth_batch(0, 2, 3.0, ragged=False)
    ref = _oracle(enc, args, None, wavs, None)
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda()
    out = enc(wav.cuda(), lens.cuda(), None, None, None)
    torch.cuda.synchronize()
    x = out["encoder_out"][0].cpu()
    assert (x - ref["encoder_out"][0]).abs().max().item() < TOL
    assert not out["encoder_padding_mask"][0].any()


def test_modality_dropout_glue(cuda):
    """Training-mode modality dropout (per-batch draws, image-drop branch) with element dropouts at 0."""
    import numpy as np
    from mm_s2ut_b200 import synth

    enc, args, cfg = _build("small")
    cfg["modality_dropout"], cfg["audio_dropout"] = 0.5, -0.5
    enc.modality_dropout, enc.audio_dropout = 0.5, -0.5
    enc.dropout_p = enc.activation_dropout_p = enc.attention_dropout_p = enc.SA_image_dropout = enc.SA_attention_dropout = 0.0
    wavs, _ = synth.synth_batch(0, 2, 3.0, ragged=True)
    imgs = synth.synth_images(0, 2)
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda().train()
    for seed in (0, 1, 2, 3):
        rs = np.random.RandomState(seed)
        draws = tuple(np.random.RandomState(seed).random_sample(2))
        enc.modality_rng = rs
        out = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[imgs.cuda()], img_masks_list=[None])
        torch.cuda.synchronize()
        ref = _oracle(enc, args, cfg, wavs, imgs, training=True, draws=draws)
        assert _compare(out, ref) < TOL


def test_base_config1_shape_parity(cuda):
    """BASELINE configs[1] architecture (12 layers, d=512, 8 heads) at B=8 x 10 s: fused states within 2e-2."""
    from mm_s2ut_b200 import synth

    enc, args, cfg = _build("base")
    wavs, _ = synth.synth_batch(1, 8, 10.0, ragged=True, zero_utt=7)
    imgs = synth.synth_images(1, 8)
    ref = _oracle(enc, args, cfg, wavs, imgs)
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda()
    out = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[imgs.cuda()], img_masks_list=[None])
    torch.cuda.synchronize()
    err = _compare(out, ref)
    record("configs[1] base B=8x10s ragged+zero utt, selective_attention: fused states max-abs err", err, TOL)
    assert err < TOL, err


def test_fp16_operands_tighter(cuda):
    """Same kernels with fp16 operands (one template switch): an order of magnitude closer to the fp32 oracle."""
    from mm_s2ut_b200 import synth

    enc, args, cfg = _build("small", op_dtype=torch.float16)
    wavs, _ = synth.synth_batch(0, 4, 5.0, ragged=True)
    imgs = synth.synth_images(0, 4)
    ref = _oracle(enc, args, cfg, wavs, imgs)
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda()
    out = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[imgs.cuda()], img_masks_list=[None])
    torch.cuda.synchronize()
    err = _compare(out, ref)
    record("configs[0] small, fp16 operands: fused states max-abs err", err, 4e-3)
    assert err < 4e-3


def test_decoder_unit_argmax_agreement(cuda):
    """north_star: >= 99 % agreement of the S2UT decoder's unit arg-max when fed the CUDA path's fused states vs the
    fp32 oracle's (same restated fairseq TransformerUnitDecoder, teacher forced on seeded units)."""
    from mm_s2ut_b200 import synth
    from oracle import decoder as odec

    enc, args, cfg = _build("base")
    wavs, _ = synth.synth_batch(1, 6, 6.0, ragged=True)
    imgs = synth.synth_images(1, 6)
    ref = _oracle(enc, args, cfg, wavs, imgs)
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda()
    out = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[imgs.cuda()], img_masks_list=[None])
    torch.cuda.synchronize()
    dsd = odec.init_decoder(args.decoder_embed_dim, args.decoder_ffn_embed_dim, args.decoder_layers, seed=3)
    g = torch.Generator().manual_seed(11)
    prev = torch.randint(4, 1004, (6, 80), generator=g)
    prev[:, 0] = 2
    mask = ref["encoder_padding_mask"][0]
    with torch.no_grad():
        l_ref = odec.unit_decoder_forward(dsd, prev, ref["encoder_out"][0], mask, args.decoder_attention_heads)
        l_gpu = odec.unit_decoder_forward(dsd, prev, out["encoder_out"][0].float().cpu(), mask,
                                          args.decoder_attention_heads)
    agree = (l_ref.argmax(-1) == l_gpu.argmax(-1)).float().mean().item()
    record("S2UT decoder unit arg-max agreement (base, 6 utt x 80 units, teacher forced)", agree, 0.99)
    assert agree >= 0.99, agree
