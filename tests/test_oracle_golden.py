"""CPU: the oracle restatements against the committed golden vectors (tests/golden/, made by oracle/make_golden.py
from the reference's own fuse.py, real torchaudio and HF's port of the fairseq S2T encoder)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from _util import fbank_errors

G = Path(__file__).parent / "golden"


def _sd(z, prefix=""):
    return {prefix + k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}


def test_selective_attention_matches_reference_fuse_py():
    from oracle import fusion

    z = np.load(G / "fuse_selective_attention.npz")
    sd = _sd(z, "sa.")
    q, img, mask = torch.from_numpy(z["q"]), torch.from_numpy(z["img"]), torch.from_numpy(z["mask"])
    out, attn = fusion.selective_attention(sd, "sa.", q, img, img, None)
    assert torch.allclose(out, torch.from_numpy(z["out"]), atol=2e-6, rtol=1e-5)
    assert torch.allclose(attn, torch.from_numpy(z["attn"]), atol=1e-7, rtol=1e-5)
    out, attn = fusion.selective_attention(sd, "sa.", q, img, img, mask)
    assert torch.allclose(out, torch.from_numpy(z["out_masked"]), atol=2e-6, rtol=1e-5)
    assert torch.allclose(attn, torch.from_numpy(z["attn_masked"]), atol=1e-7, rtol=1e-5)
    assert attn[1, :, 40:].abs().max() == 0


def test_multimodal_attention_matches_reference_fuse_py():
    from oracle import fusion

    z = np.load(G / "fuse_multimodal_attention.npz")
    sd = _sd(z, "ma.")
    q, img, mask = torch.from_numpy(z["q"]), torch.from_numpy(z["img"]), torch.from_numpy(z["mask"])
    out = fusion.multimodal_attention(sd, "ma.", q, img, None)
    assert torch.allclose(out, torch.from_numpy(z["out"]), atol=2e-6, rtol=1e-5)
    out = fusion.multimodal_attention(sd, "ma.", q, img, mask)
    assert torch.allclose(out, torch.from_numpy(z["out_masked"]), atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("case", ["sa_gate_prenorm", "sa_nogate", "ma_gate_prenorm", "sa_two_types"])
def test_fuse_img_feat_and_glue_match_reference_statements(case):
    """oracle.fusion.fuse_img_feat / fusion_top against the reference's OWN fuse_img_feat method and the
    fusion-at-top statements of its forward() (mm_s2s_transformer.py:594-622, :496-560), which
    oracle/make_golden.py cut out with ``ast`` and ran bound to the reference's fuse.py modules."""
    from types import SimpleNamespace

    from oracle import fusion

    z = np.load(G / f"fuse_img_feat_{case}.npz")
    sd = _sd(z)
    n = int(z["n_types"])
    cfg = SimpleNamespace(image_pre_norm=bool(z["pre_norm"]), multimodal_attention_type=str(z["kind"]),
                          use_selective_gate=bool(z["gate"]), modality_dropout=0.5, audio_dropout=-0.5,
                          is_fusion_top=True)
    text = torch.from_numpy(z["text"])
    imgs = [torch.from_numpy(z[f"img{j}"]) for j in range(n)]
    masks = [torch.from_numpy(z[f"mask{j}"]) for j in range(n)]
    tol = dict(atol=3e-6, rtol=1e-5)
    res = fusion.fuse_img_feat(sd, "", cfg, text, 0, imgs[0].transpose(0, 1), None)
    assert torch.allclose(res, torch.from_numpy(z["res"]), **tol)
    res = fusion.fuse_img_feat(sd, "", cfg, text, 0, imgs[0].transpose(0, 1), masks[0])
    assert torch.allclose(res, torch.from_numpy(z["res_masked"]), **tol)
    for tag, training, draws in [("eval", False, None), ("keep", True, (0.9, 0.9)), ("drop_image", True, (0.1, 0.9))]:
        got = fusion.fusion_top(sd, "", cfg, text, imgs, masks, training, draws)
        assert torch.allclose(got, torch.from_numpy(z[f"glue_{tag}"]), **tol), tag
    assert not np.allclose(z["glue_keep"], z["glue_drop_image"])      # the image-drop branch did something
    assert np.array_equal(z["glue_keep"], z["glue_eval"])


def test_fbank_restatement_matches_torchaudio_golden():
    from oracle import fbank as ofb

    z = np.load(G / "fbank_torchaudio.npz")
    for u in range(3):
        got = ofb.kaldi_fbank_np(z[f"wav{u}"])
        ref = z[f"fbank{u}"]
        assert got.shape == ref.shape
        e_main, e_all = fbank_errors(got, ref)
        assert e_main < 1e-4 and e_all < 5e-4, (e_main, e_all)   # measured: 3e-6 .. 5e-6 / 8e-6 .. 3.1e-4
        live = ofb.kaldi_fbank_ta(z[f"wav{u}"])         # the installed torchaudio still agrees with the fixture
        assert np.max(np.abs(live - ref)) < 1e-4
    assert np.array_equal(ofb.kaldi_fbank_np(z["wav_zero"]), z["fbank_zero"])   # silence: the log(eps) floor


def test_s2t_encoder_restatement_matches_hf_port():
    from oracle import s2t

    z = np.load(G / "hf_speech2text_encoder.npz")
    sd = _sd(z)
    out = s2t.s2t_encoder_forward(sd, torch.from_numpy(z["feats"]), torch.from_numpy(z["lens"]), int(z["heads"]))
    x = out["encoder_out"][0].transpose(0, 1)       # [B, T', d]
    ref = torch.from_numpy(z["out"])
    mask = out["encoder_padding_mask"][0]
    assert x.shape == ref.shape
    err = ((x - ref).abs() * (~mask).unsqueeze(-1)).max().item()
    assert err < 1e-4, err


def test_unit_decoder_restatement_matches_hf_port():
    """oracle/decoder.py (fairseq TransformerDecoder as the S2UT unit decoder uses it) against HF's Speech2TextDecoder,
    the port of the same fairseq module: scaled embedding, fairseq positions with right-padded targets, causal
    self-attention, encoder attention under an encoder padding mask, tied output projection."""
    from oracle import decoder as od

    z = np.load(G / "hf_speech2text_decoder.npz")
    sd = _sd(z)
    tokens, enc, lens = torch.from_numpy(z["tokens"]), torch.from_numpy(z["enc"]), torch.from_numpy(z["enc_lens"])
    mask = torch.arange(enc.shape[1])[None, :] >= lens[:, None]
    out = od.unit_decoder_forward(sd, tokens, enc.transpose(0, 1).contiguous(), mask, int(z["heads"]))
    ref = torch.from_numpy(z["logits"])
    assert out.shape == ref.shape and ref.abs().max().item() > 1.0
    err = ((out - ref).abs() * tokens.ne(1).unsqueeze(-1)).max().item()
    assert err < 1e-4, err


def test_adam_restatement_matches_torch_adam_in_the_eps_to_zero_limit():
    """fairseq's Adam (oracle/adam.py) and torch.optim.Adam differ only in where eps enters (sqrt(v) + eps vs
    sqrt(v / bias_correction2) + eps): with a vanishing eps the two must walk the same trajectory."""
    from oracle import adam as oadam

    rng = np.random.default_rng(3)
    p0 = rng.standard_normal(257).astype(np.float32)
    grads = [rng.standard_normal(257).astype(np.float32) * s for s in (1.0, 0.3, 2.0, 0.05, 1.0)]
    t = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([t], lr=5e-4, betas=(0.9, 0.98), eps=1e-30)
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for i, g in enumerate(grads):
        t.grad = torch.from_numpy(g.copy())
        opt.step()
        p, m, v = oadam.adam_step(p, g, m, v, lr=5e-4, betas=(0.9, 0.98), eps=1e-30, step=i + 1)
        assert np.max(np.abs(p - t.detach().numpy())) < 2e-7 * (i + 1)
    assert np.max(np.abs(p - p0)) > 1e-3          # the parameters did move


def test_utterance_cmvn_semantics():
    from oracle import fbank as ofb

    rng = np.random.default_rng(0)
    x = (rng.standard_normal((57, 80)) * 3 + 11).astype(np.float32)
    y = ofb.utterance_cmvn(x)
    assert abs(float(y.mean())) < 1e-4 and abs(float(y.std()) - 1.0) < 1e-2
    out, lens = ofb.collate_frames([y, y[:20]])
    assert out.shape == (2, 57, 80) and lens.tolist() == [57, 20] and np.all(out[1, 20:] == 0)


def _grads_match(z, sd, grads_of, atol=3e-5, rtol=2e-4):
    for k in z.files:
        if k.startswith("grad."):
            got = grads_of[k[5:]]
            ref = torch.from_numpy(z[k])
            assert torch.allclose(got, ref, atol=atol * max(1.0, ref.abs().max().item()), rtol=rtol), k


def test_selective_attention_gradients_match_reference_fuse_py_autograd():
    """The oracle's differentiable path against gradients computed by autograd through the reference's own fuse.py
    (the yardstick of the CUDA backward pass is pinned by the reference itself)."""
    from oracle import fusion

    z = np.load(G / "fuse_selective_attention_grads.npz")
    sd = {k: v.clone().requires_grad_() for k, v in _sd(z, "sa.").items()}
    q, img = torch.from_numpy(z["q"]).requires_grad_(), torch.from_numpy(z["img"]).requires_grad_()
    out, _ = fusion.selective_attention(sd, "sa.", q, img, img, None)
    (out * torch.from_numpy(z["R"])).sum().backward()
    _grads_match(z, sd, {k[3:]: v.grad for k, v in sd.items()})
    assert torch.allclose(q.grad, torch.from_numpy(z["dq"]), atol=1e-5, rtol=1e-4)
    assert torch.allclose(img.grad, torch.from_numpy(z["dimg"]), atol=1e-5, rtol=1e-4)


def test_multimodal_attention_gradients_match_reference_fuse_py_autograd():
    from oracle import fusion

    z = np.load(G / "fuse_multimodal_attention_grads.npz")
    sd = {k: v.clone().requires_grad_() for k, v in _sd(z, "ma.").items()}
    q, img = torch.from_numpy(z["q"]).requires_grad_(), torch.from_numpy(z["img"]).requires_grad_()
    out = fusion.multimodal_attention(sd, "ma.", q, img, None)
    (out * torch.from_numpy(z["R"])).sum().backward()
    _grads_match(z, sd, {k[3:]: v.grad for k, v in sd.items()})
    assert torch.allclose(q.grad, torch.from_numpy(z["dq"]), atol=1e-5, rtol=1e-4)
    assert torch.allclose(img.grad, torch.from_numpy(z["dimg"]), atol=1e-5, rtol=1e-4)
