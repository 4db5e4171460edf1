"""Host logic of the inference forward, checked on CPU against the fp32 oracle with the kernel wrappers emulated
(``tests/_emul.py``): buffer layout of the fused speech -> image attention path (one [B, Tk, 2d] K|V tensor, the
learned bias_k | bias_v row, the key mask) and of the scores / softmax / P V path it replaces."""
import pytest
import torch

import _emul


def _emulated(monkeypatch):
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200 import engine, training

    monkeypatch.setattr(engine, "K", _emul)
    monkeypatch.setattr(training, "K", _emul)
    monkeypatch.setattr(engine.EncoderEngine, "_require_cuda", False)


@pytest.mark.parametrize("attn_type", ["selective_attention", "multimodal_attention"])
@pytest.mark.parametrize("fused", [True, False])
def test_eval_forward_fused_and_unfused_attention_match_oracle(monkeypatch, attn_type, fused):
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder
    from oracle import fbank as ofb, fusion as ofu

    _emulated(monkeypatch)
    cfg = dict(load_mm_config(DEFAULT_YAML))
    cfg["multimodal_attention_type"] = attn_type
    torch.manual_seed(1)
    args = make_args("small", multimodal_translation_config_yaml=cfg)
    enc = MM_S2STransformerEncoder(args, build_unused_projections=False).eval()
    enc.fuse_cross_attention = fused
    sd = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    wavs, _ = synth.synth_batch(2, 2, 1.0, ragged=True)
    imgs = synth.synth_images(2, 2, 50, 768)
    mask = torch.zeros(2, 50, dtype=torch.bool)
    mask[1, 30:] = True
    feats, flens = ofb.features_from_waveforms(wavs)
    wav, lens = synth.pad_waveforms(wavs)
    for m in (None, mask):
        ref = ofu.mm_encoder_forward(sd, load_mm_config(cfg), torch.from_numpy(feats), torch.from_numpy(flens), [imgs],
                                     [m], args.encoder_attention_heads)
        out = enc(wav, lens, None, None, None, imgs_list=[imgs], img_masks_list=[m])
        assert enc.engine().fused_xattn is fused
        valid = (~ref["encoder_padding_mask"][0]).t().unsqueeze(-1)
        err = ((out["encoder_out"][0] - ref["encoder_out"][0]).abs() * valid).max().item()
        assert err < 2e-2, (attn_type, fused, m is not None, err)
