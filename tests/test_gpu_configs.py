"""Parity on the other BASELINE configurations and edge shapes (all through the module API / C ABI)."""
import pytest
import torch

from _util import record

from test_gpu_encoder import TOL, _build, _compare, _oracle

pytestmark = pytest.mark.gpu


def test_large_config4_detr_feats(cuda):
    """BASELINE configs[4] architecture: 16 layers, d=1024, 16 heads, DETR-style 100 x 256 image features."""
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder

    cfg = dict(load_mm_config(DEFAULT_YAML))
    cfg["image_feat_dim"] = [256]
    torch.manual_seed(4)
    args = make_args("large", multimodal_translation_config_yaml=cfg)
    enc = MM_S2STransformerEncoder(args, build_unused_projections=False).eval()
    wavs, _ = synth.synth_batch(4, 3, 4.0, ragged=True)
    imgs = synth.synth_images(4, 3, 100, 256)
    ref = _oracle(enc, args, load_mm_config(cfg), wavs, imgs)
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda()
    out = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[imgs.cuda()], img_masks_list=[None])
    torch.cuda.synchronize()
    err = _compare(out, ref)
    record("configs[4] large (16 layers, d=1024) B=3x4s, 100x256 image feats: fused states max-abs err", err, TOL)
    assert err < TOL, err


def test_long_utterances_chunked_attention(cuda):
    """15 s and 28 s utterances: T = 375 / 700 > 256 exercises the chunked two-sweep attention kernel in the encoder."""
    from mm_s2ut_b200 import synth

    enc, args, cfg = _build("small")
    wavs = [synth.synth_waveform(3, 0, 28.0, ragged=False), synth.synth_waveform(3, 1, 15.0, ragged=False)]
    imgs = synth.synth_images(3, 2)
    ref = _oracle(enc, args, cfg, wavs, imgs)
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda()
    out = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[imgs.cuda()], img_masks_list=[None])
    torch.cuda.synchronize()
    assert out["encoder_out"][0].shape[0] == 700
    err = _compare(out, ref)
    assert err < TOL, err


def test_two_image_feature_types_sum(cuda):
    """image_feat_dim: [256, 768] -> one attention + gate per type, fused states are the SUM over types
    (mm_s2s_transformer.py:557-560).  image_pre_norm off: the reference's single shared LayerNorm cannot serve two dims."""
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder
    from oracle import fbank as ofb, fusion as ofu

    cfg = dict(load_mm_config(DEFAULT_YAML))
    cfg["image_feat_dim"] = [256, 768]
    cfg["image_pre_norm"] = False
    torch.manual_seed(5)
    args = make_args("small", multimodal_translation_config_yaml=cfg)
    enc = MM_S2STransformerEncoder(args, build_unused_projections=False).eval()
    wavs, _ = synth.synth_batch(5, 3, 3.0, ragged=True)
    img_a, img_b = synth.synth_images(5, 3, 100, 256), synth.synth_images(6, 3, 577, 768)
    sd = {k: v.detach() for k, v in enc.state_dict().items()}
    feats, flens = ofb.features_from_waveforms(wavs)
    ref = ofu.mm_encoder_forward(sd, load_mm_config(cfg), torch.from_numpy(feats), torch.from_numpy(flens),
                                 [img_a, img_b], [None, None], args.encoder_attention_heads)
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda()
    out = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[img_a.cuda(), img_b.cuda()],
              img_masks_list=[None, None])
    torch.cuda.synchronize()
    assert _compare(out, ref) < 1.5 * TOL      # the sum of two fused streams has twice the magnitude of one


def test_batch_one_and_short_utterance(cuda):
    """B = 1 and a 0.5 s utterance (T = 13): tiles mostly empty, everything masked by TMA bounds."""
    from mm_s2ut_b200 import synth

    enc, args, cfg = _build("small")
    wavs = [synth.synth_waveform(9, 0, 0.5, ragged=False)]
    imgs = synth.synth_images(9, 1)
    ref = _oracle(enc, args, cfg, wavs, imgs)
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda()
    out = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[imgs.cuda()], img_masks_list=[None])
    torch.cuda.synchronize()
    x, r = out["encoder_out"][0].cpu(), ref["encoder_out"][0]
    assert x.shape == r.shape and (x - r).abs().max().item() < TOL


def test_full_size_config1_properties(cuda):
    """BASELINE configs[1] at full size (B=64 x 10 s): size-independent properties instead of a CPU oracle run --
    (1) batch independence: utterance i encoded inside the 64-batch equals the same utterance encoded in a batch of 4
    bit-for-bit is not required (tile boundaries differ) but must agree to 1e-2; (2) determinism: two runs are bit-equal;
    (3) padded positions never influence valid ones (perturbing the padding leaves valid outputs bit-equal)."""
    from mm_s2ut_b200 import synth

    enc, args, cfg = _build("base")
    enc.cuda()
    wavs, _ = synth.synth_batch(1, 64, 10.0, ragged=True)
    imgs = synth.synth_images(1, 64).cuda()
    wav, lens = synth.pad_waveforms(wavs)
    wav, lens = wav.cuda(), lens.cuda()
    o1 = enc(wav, lens, None, None, None, imgs_list=[imgs], img_masks_list=[None])["encoder_out"][0].clone()
    o2 = enc(wav, lens, None, None, None, imgs_list=[imgs], img_masks_list=[None])["encoder_out"][0].clone()
    assert torch.equal(o1, o2)
    wav2 = wav.clone()
    for b in range(64):
        wav2[b, int(lens[b]):] = 12345.0          # garbage in the sample padding
    o3 = enc(wav2, lens, None, None, None, imgs_list=[imgs], img_masks_list=[None])
    mask = o3["encoder_padding_mask"][0]
    valid = (~mask).t().unsqueeze(-1)
    assert torch.equal(o1 * valid, o3["encoder_out"][0] * valid)
    idx = [0, 17, 40, 63]
    sub_wav, sub_len = synth.pad_waveforms([wavs[i] for i in idx])
    o4 = enc(sub_wav.cuda(), sub_len.cuda(), None, None, None, imgs_list=[imgs[idx]], img_masks_list=[None])
    T4 = o4["encoder_out"][0].shape[0]
    for j, i in enumerate(idx):
        # the last conv window of an utterance sees either zero FRAMES (longer batch: conv1 of zeros = GLU(bias))
        # or the conv's own zero padding (utterance is the longest of its batch) -- fairseq semantics, so the
        # final two positions legitimately depend on the batch; everything before them must agree
        n = int((~o4["encoder_padding_mask"][0][j]).sum()) - 2
        assert (o1[:n, i] - o4["encoder_out"][0][:n, j]).abs().max().item() < 1e-2, (i, T4)


@pytest.mark.parametrize("ragged", [False, True])
def test_timed_configuration_graphs_vs_oracle(cuda, ragged):
    """The objects bench.py times -- ``GraphedEncoder`` replays at BASELINE configs[1]'s full size (base model,
    B = 64 x 10 s, 577 x 768 image features) with fp32-waveform, int16-PCM, fp16-image and feature-store inputs --
    against the fp32 CPU oracle on the same inputs (north-star tolerance: 2e-2 max-abs over valid positions).
    ragged=False is exactly the bench shape (every utterance 160 000 samples, no padding)."""
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.feature_store import ImageFeatureStore
    from mm_s2ut_b200.graph import GraphedEncoder

    B, n_max = 64, 160000
    enc, args, cfg = _build("base")
    wavs, _ = synth.synth_batch(1, B, 10.0, ragged=ragged, zero_utt=5 if ragged else None)
    wavs = [w.round().clip(-32768, 32767) for w in wavs]          # 16-bit PCM values: the fp32 and int16 inputs agree
    imgs = synth.synth_images(1, B)
    # the oracle sees the batch at the graph's static shape (feature rows zero-padded to the 998 frames of 10 s): in
    # fairseq the last two positions of an utterance depend on how far its batch is padded (conv of zero frames =
    # GLU(bias), not the conv's own zero padding), so both sides must be padded alike
    from oracle import fbank as ofb, fusion as ofu

    feats, flens = ofb.features_from_waveforms(wavs)
    m_max = 1 + (n_max - 400) // 160
    feats = torch.nn.functional.pad(torch.from_numpy(feats), (0, 0, 0, m_max - feats.shape[1]))
    sd = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    ref = ofu.mm_encoder_forward(sd, cfg, feats, torch.from_numpy(flens), [imgs], [None], args.encoder_attention_heads)
    if not ref["encoder_padding_mask"]:          # the oracle keeps fairseq's "no padding -> empty list"
        ref["encoder_padding_mask"] = [torch.zeros(B, ref["encoder_out"][0].shape[0], dtype=torch.bool)]
    wav, lens = synth.pad_waveforms(wavs)
    if wav.shape[1] < n_max:
        wav = torch.nn.functional.pad(wav, (0, n_max - wav.shape[1]))
    enc.cuda()
    wav, lens, imgs = wav.cuda(), lens.cuda(), imgs.cuda()
    Tref = ref["encoder_out"][0].shape[0]
    assert Tref == 250

    def check(name, ge, inputs, tol=TOL):
        ge.load_inputs(*inputs)
        ge.capture()
        for _ in range(2):                  # the replay, not the capture pass, is what the bench times
            ge.load_inputs(*inputs)
            out = ge.replay()
        torch.cuda.synchronize()
        err = _compare(out, ref)
        record(f"configs[1] FULL SIZE base B=64x10s {'ragged' if ragged else 'bench shape'}, GraphedEncoder {name}: "
               f"fused states max-abs err", err, tol)
        assert err < tol, (name, err)

    shapes = [(577, 768)]
    check("fp32 waveform", GraphedEncoder(enc, B, n_max, shapes), (wav, lens, [imgs]))
    check("int16 PCM", GraphedEncoder(enc, B, n_max, shapes, wav_dtype=torch.int16), (wav.to(torch.int16), lens, [imgs]))
    check("int16 PCM + fp16 image features", GraphedEncoder(enc, B, n_max, shapes, wav_dtype=torch.int16,
                                                            img_dtype=torch.float16),
          (wav.to(torch.int16), lens, [imgs.half()]))
    store = ImageFeatureStore(torch.cat([imgs.flip(0), imgs], 0), cuda)
    idx = torch.arange(B, 2 * B, dtype=torch.int64, device=cuda)
    check("int16 PCM + device feature store", GraphedEncoder(enc, B, n_max, shapes, wav_dtype=torch.int16, stores=[store]),
          (wav.to(torch.int16), lens, [idx]))
