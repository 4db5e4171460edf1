"""int16 PCM ingest: bit-identical to the float32 path for integer-valued audio, through the kernel and the module."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fbank_int16_equals_float(cuda):
    from mm_s2ut_b200 import kernels as K, synth

    wavs, _ = synth.synth_batch(0, 3, 2.0, ragged=True)
    wav, lens = synth.pad_waveforms(wavs)
    wav = wav.round().clamp_(-32768, 32767)
    m = 1 + (wav.shape[1] - 400) // 160
    tab = K.fbank_tables(cuda)
    f32 = torch.zeros(3, m, 80, device=cuda)
    i16 = torch.zeros(3, m, 80, device=cuda)
    K.fbank(wav.to(cuda), lens.to(cuda), f32, tab)
    K.fbank(wav.to(torch.int16).to(cuda), lens.to(cuda), i16, tab)
    torch.cuda.synchronize()
    for b in range(3):
        n = 1 + (int(lens[b]) - 400) // 160
        assert torch.equal(f32[b, :n], i16[b, :n])


def test_encoder_accepts_int16_waveform(cuda):
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.config import DEFAULT_YAML, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder

    torch.manual_seed(0)
    enc = MM_S2STransformerEncoder(make_args("small", multimodal_translation_config_yaml=str(DEFAULT_YAML)),
                                   build_unused_projections=False).eval().cuda()
    wavs, _ = synth.synth_batch(0, 2, 2.0, ragged=True)
    wav, lens = synth.pad_waveforms(wavs)
    wav = wav.round().clamp_(-32768, 32767)
    imgs = synth.synth_images(0, 2).cuda()
    a = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[imgs], img_masks_list=[None])["encoder_out"][0].clone()
    b = enc(wav.to(torch.int16).cuda(), lens.cuda(), None, None, None, imgs_list=[imgs],
            img_masks_list=[None])["encoder_out"][0]
    torch.cuda.synchronize()
    assert torch.equal(a, b)
