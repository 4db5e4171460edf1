"""S2UT unit decoder (teacher forced) on the CUDA kernels vs the oracle's fp32 restatement, fed the same encoder
states: logits close, unit arg-max agreement >= 99 % (north_star)."""
import pytest
import torch

from _util import record

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,ffn,heads,layers,L,T", [(512, 2048, 8, 6, 80, 150), (256, 1024, 4, 2, 300, 40)])
def test_unit_decoder_forward(cuda, d, ffn, heads, layers, L, T):
    from mm_s2ut_b200.decoder import UnitDecoderEngine
    from oracle import decoder as odec

    B = 3
    sd = odec.init_decoder(d, ffn, layers, seed=5)
    g = torch.Generator().manual_seed(d + L)
    prev = torch.randint(4, 1004, (B, L), generator=g)
    prev[:, 0] = 2
    enc = torch.randn(T, B, d, generator=g) * 0.6
    lens = torch.tensor([T, T - 17, max(1, T // 3)])
    mask = torch.arange(T)[None, :] >= lens[:, None]
    with torch.no_grad():
        ref = odec.unit_decoder_forward(sd, prev, enc, mask, heads)
    eng = UnitDecoderEngine(sd, heads, cuda)
    out = eng.forward(prev.to(cuda), enc.to(cuda), mask.to(cuda))
    torch.cuda.synchronize()
    out = out.float().cpu()
    assert out.shape == ref.shape
    err = (out - ref).abs().max().item()
    agree = (out.argmax(-1) == ref.argmax(-1)).float().mean().item()
    record(f"unit decoder d={d} layers={layers} L={L}: logits max-abs err vs fp32 oracle", err, 0.06)
    record(f"unit decoder d={d} layers={layers} L={L}: unit arg-max agreement", agree, 0.99)
    assert err < 0.06, err
    assert agree >= 0.99, agree


def test_decoder_on_gpu_encoder_states(cuda):
    """Whole chain on the GPU: waveform -> fused encoder states -> unit logits, against oracle encoder + oracle decoder."""
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.decoder import UnitDecoderEngine
    from oracle import decoder as odec
    from test_gpu_encoder import _build, _oracle

    enc, args, cfg = _build("base")
    wavs, _ = synth.synth_batch(1, 4, 5.0, ragged=True)
    imgs = synth.synth_images(1, 4)
    ref_enc = _oracle(enc, args, cfg, wavs, imgs)
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda()
    out = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[imgs.cuda()], img_masks_list=[None])
    dsd = odec.init_decoder(args.decoder_embed_dim, args.decoder_ffn_embed_dim, args.decoder_layers, seed=3)
    g = torch.Generator().manual_seed(11)
    prev = torch.randint(4, 1004, (4, 60), generator=g)
    prev[:, 0] = 2
    mask = ref_enc["encoder_padding_mask"][0]
    with torch.no_grad():
        l_ref = odec.unit_decoder_forward(dsd, prev, ref_enc["encoder_out"][0], mask, args.decoder_attention_heads)
    eng = UnitDecoderEngine(dsd, args.decoder_attention_heads, cuda)
    l_gpu = eng.forward(prev.to(cuda), out["encoder_out"][0], out["encoder_padding_mask"][0]).float().cpu()
    agree = (l_ref.argmax(-1) == l_gpu.argmax(-1)).float().mean().item()
    record("waveform -> unit logits entirely on the GPU (base, 4 utt x 60 units): arg-max agreement with the oracle chain",
           agree, 0.99)
    assert agree >= 0.99, agree


def test_label_smoothed_nll(cuda):
    """Criterion forward on the padded logits buffer (ld 1008, V 1004) with padded targets, vs the oracle in fp64-ish."""
    from mm_s2ut_b200 import kernels as K
    from oracle import decoder as odec

    g = torch.Generator().manual_seed(9)
    rows, V, ld = 3 * 157, 1004, 1008
    logits = torch.randn(rows, ld, generator=g) * 3
    target = torch.randint(4, V, (rows,), generator=g)
    target[::7] = 1                                   # padding
    ref_loss, ref_nll = odec.label_smoothed_nll_loss(logits[:, :V].double(), target, 0.2)
    loss, nll = K.label_smoothed_nll(logits.to(cuda), V, target.to(cuda), 1, 0.2)
    torch.cuda.synchronize()
    assert abs(loss.item() - ref_loss.item()) / abs(ref_loss.item()) < 2e-5
    assert abs(nll.item() - ref_nll.item()) / abs(ref_nll.item()) < 2e-5


def test_model_forward_mirrors_reference_surface(cuda):
    """MM_S2UTTransformerModel.forward: 9 encoder kwargs through, decoder on the encoder output, encoder states attached
    when return_all_hiddens (reference mm_s2s_transformer.py:667-700); logits agree with oracle encoder + oracle decoder."""
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args
    from mm_s2ut_b200.models.mm_s2ut_model import MM_S2UTTransformerModel
    from oracle import decoder as odec
    from test_gpu_encoder import _oracle

    torch.manual_seed(2)
    args = make_args("small", multimodal_translation_config_yaml=str(DEFAULT_YAML))
    model = MM_S2UTTransformerModel(args, build_unused_projections=False).eval()
    keys = set(model.state_dict().keys())
    assert "decoder.embed_tokens.weight" in keys and "decoder.output_projection.weight" in keys
    assert "decoder.layers.0.encoder_attn.k_proj.weight" in keys and "encoder.layer_norm.weight" in keys
    wavs, _ = synth.synth_batch(0, 3, 4.0, ragged=True)
    imgs = synth.synth_images(0, 3)
    ref_enc = _oracle(model.encoder, args, load_mm_config(DEFAULT_YAML), wavs, imgs)
    dsd = {k: v.detach().clone() for k, v in model.decoder.state_dict().items() if not k.startswith("output_projection.")}
    g = torch.Generator().manual_seed(4)
    prev = torch.randint(4, 1004, (3, 40), generator=g)
    prev[:, 0] = 2
    with torch.no_grad():
        l_ref = odec.unit_decoder_forward(dsd, prev, ref_enc["encoder_out"][0], ref_enc["encoder_padding_mask"][0],
                                          args.decoder_attention_heads)
    wav, lens = synth.pad_waveforms(wavs)
    model.cuda()
    logits, extra = model(wav.cuda(), lens.cuda(), prev.cuda(), None, None, None, imgs_list=[imgs.cuda()],
                          img_masks_list=[None], return_all_hiddens=True)
    torch.cuda.synchronize()
    assert logits.shape == l_ref.shape
    assert len(extra["encoder_states"]) == args.encoder_layers and len(extra["encoder_padding_mask"]) == 1
    agree = (logits.float().cpu().argmax(-1) == l_ref.argmax(-1)).float().mean().item()
    assert agree >= 0.99, agree
