"""Row-wise kernels (LayerNorm, softmax, convert, CMVN, fbank) and the self-attention core vs references."""
import numpy as np
import pytest
import torch

from _util import fbank_errors, record

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim", [256, 512, 768, 1024])
def test_layernorm(cuda, dim):
    from mm_s2ut_b200 import kernels as K

    g = torch.Generator().manual_seed(dim)
    x = (torch.randn(1003, dim, generator=g) * 3 + 0.7).to(cuda)
    gamma, beta = torch.randn(dim, generator=g).to(cuda), torch.randn(dim, generator=g).to(cuda)
    o16 = torch.zeros(1003, dim, dtype=torch.bfloat16, device=cuda)
    o32 = torch.zeros(1003, dim, dtype=torch.float32, device=cuda)
    K.layernorm(x, gamma, beta, out_op=o16, out_f32=o32)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x, (dim,), gamma, beta, 1e-5)
    assert (o32 - ref).abs().max().item() < 2e-5
    assert (o16.float() - ref).abs().max().item() < 4e-2


@pytest.mark.parametrize("Tk,Tkp", [(577, 584), (577, 579), (50, 52), (1000, 1024), (197, 200)])
def test_softmax_rows(cuda, Tk, Tkp):
    """Aligned leading dimensions take the 128-bit kernel, 579 the scalar one; pad columns hold NaN on input."""
    from mm_s2ut_b200 import kernels as K

    g = torch.Generator().manual_seed(5)
    B, T = 3, 50
    S = (torch.randn(B * T, Tkp, generator=g) * 4).to(cuda)
    S[:, Tk:] = float("nan")
    mask = (torch.rand(B, Tk, generator=g) < 0.2).to(torch.uint8).to(cuda)
    P = torch.full((B * T, Tkp), 3.0, dtype=torch.bfloat16, device=cuda)
    K.softmax_rows(S, Tkp, B * T, Tk, P, Tkp, key_mask=mask, rows_per_seq=T)
    torch.cuda.synchronize()
    s = S[:, :Tk].view(B, T, Tk).masked_fill(mask.bool()[:, None, :], float("-inf"))
    ref = torch.softmax(s, -1).view(B * T, Tk)
    assert (P[:, :Tk].float() - ref).abs().max().item() < 2e-3
    assert P[:, Tk:].abs().max().item() == 0


def test_convert(cuda):
    from mm_s2ut_b200 import kernels as K

    x = torch.randn(12345, device=cuda)[:12343].clone()
    for dt in (torch.bfloat16, torch.float16):
        o = torch.zeros(12343, dtype=dt, device=cuda)
        K.convert(x, o)
        torch.cuda.synchronize()
        assert torch.equal(o, x.to(dt))


def _fbank_batch(cfg, n, dur, zero_utt=None):
    from mm_s2ut_b200 import synth

    wavs, lens = synth.synth_batch(cfg, n, dur, ragged=True, zero_utt=zero_utt)
    return wavs, synth.pad_waveforms(wavs)


def test_fbank_vs_torchaudio(cuda):
    """fbank (pre-CMVN) within 1e-4 relative of torchaudio.compliance.kaldi.fbank in fp32 (north_star tolerance)."""
    from mm_s2ut_b200 import kernels as K
    from oracle import fbank as ofb

    wavs, (wav, lens) = _fbank_batch(0, 6, 3.0, zero_utt=5)
    wav, lens = wav.to(cuda), lens.to(cuda)
    m = 1 + (wav.shape[1] - 400) // 160
    feats = torch.zeros(len(wavs), m, 80, device=cuda)
    K.fbank(wav, lens, feats, K.fbank_tables(cuda))
    torch.cuda.synchronize()
    for i, w in enumerate(wavs):
        ref = torch.from_numpy(ofb.kaldi_fbank_ta(w))
        got = feats[i, : ref.shape[0]].cpu()
        e_main, e_all = fbank_errors(got.numpy(), ref.numpy())
        record(f"fbank vs torchaudio, utterance {i}: max rel err (bins within 60 dB of frame max)", e_main, 1e-4)
        record(f"fbank vs torchaudio, utterance {i}: max rel err (all bins)", e_all, 2e-4)
        assert e_main < 1e-4 and e_all < 2e-4, (i, e_main, e_all)
    # silence: every bin sits on the log(eps) floor, bit-identical to torch's value
    assert torch.equal(feats[5, : 1 + (len(wavs[5]) - 400) // 160].cpu(), torch.from_numpy(ofb.kaldi_fbank_ta(wavs[5])))


def test_fbank_edge_lengths_and_unaligned_rows(cuda):
    """Lengths the chunking has to get right -- shorter than one window (no frame, nothing written), exactly one window,
    one short of / one past a 24-frame work item, one past an 8-frame round -- and a row stride that leaves every
    utterance but the first off the 16-byte grid (the scalar staging path), fp32 and int16 alike: each utterance's
    frames are the same bits as when it is run alone from an aligned buffer, and rows past the last frame stay untouched."""
    from mm_s2ut_b200 import kernels as K

    g = torch.Generator().manual_seed(11)
    n_list = [399, 400, 559, 560, 400 + 160 * 22, 400 + 160 * 23, 400 + 160 * 24, 400 + 160 * 8, 400 + 160 * 47 + 159]
    stride = max(n_list) + 3                                    # odd stride: rows 1.. start off the 16-byte grid
    wav = torch.zeros(len(n_list), stride)
    for i, n in enumerate(n_list):
        wav[i, :n] = (torch.randn(n, generator=g) * 4000).round().clamp(-32768, 32767)
    lens = torch.tensor(n_list, dtype=torch.int64)
    m = 1 + (max(n_list) - 400) // 160
    tables = K.fbank_tables(cuda)
    for dt in (torch.float32, torch.int16):
        feats = torch.full((len(n_list), m, 80), 7.0, device=cuda)
        K.fbank(wav.to(dt).to(cuda), lens.to(cuda), feats, tables)
        torch.cuda.synchronize()
        for i, n in enumerate(n_list):
            nf = 0 if n < 400 else 1 + (n - 400) // 160
            assert (feats[i, nf:] == 7.0).all(), (dt, n)        # nothing written past the utterance's frames
            if nf == 0:
                continue
            alone = torch.full((1, nf, 80), 7.0, device=cuda)
            K.fbank(wav[i: i + 1, : n + (-n) % 8].contiguous().to(dt).to(cuda), lens[i: i + 1].to(cuda), alone, tables)
            torch.cuda.synchronize()
            assert torch.isfinite(alone).all()
            assert torch.equal(feats[i, :nf], alone[0]), (dt, n)


def test_fbank_full_size_shift_property(cuda):
    """BASELINE size (64 x 10 s): a size-independent property instead of the CPU oracle.  Frame f of an utterance depends
    on samples [160 f, 160 f + 400) only, so the features of the waveform advanced by 160 k samples are rows k.. of the
    original's, bit for bit (same samples, same arithmetic, whichever work item, half-warp and round they land in); and
    an utterance's features do not depend on its position in the batch."""
    from mm_s2ut_b200 import kernels as K

    g = torch.Generator().manual_seed(12)
    B, n, k = 64, 160000, 37
    wav = (torch.randn(B, n, generator=g) * 3000).round().to(cuda)
    lens = torch.full((B,), n, dtype=torch.int64, device=cuda)
    m = 1 + (n - 400) // 160
    tables = K.fbank_tables(cuda)
    feats = torch.zeros(B, m, 80, device=cuda)
    K.fbank(wav, lens, feats, tables)
    shifted = wav[:, 160 * k:].contiguous()
    lens_s = torch.full((B,), n - 160 * k, dtype=torch.int64, device=cuda)
    feats_s = torch.zeros(B, m - k, 80, device=cuda)
    K.fbank(shifted, lens_s, feats_s, tables)
    perm = torch.randperm(B, generator=g).to(cuda)
    feats_p = torch.zeros(B, m, 80, device=cuda)
    K.fbank(wav[perm].contiguous(), lens, feats_p, tables)
    torch.cuda.synchronize()
    assert torch.isfinite(feats).all()
    assert torch.equal(feats[:, k:], feats_s)
    assert torch.equal(feats[perm], feats_p)


def test_cmvn_matches_reference_arithmetic(cuda):
    """CMVN statistics replay numpy's sequential fp32 accumulation: given the SAME fbank values the normalised
    features are bit-identical to fairseq UtteranceCMVN (incl. the all-zero utterance's floor behaviour)."""
    from mm_s2ut_b200 import kernels as K
    from oracle import fbank as ofb

    wavs, (wav, lens) = _fbank_batch(0, 4, 2.0, zero_utt=3)
    B = len(wavs)
    raw = [ofb.kaldi_fbank_ta(w) for w in wavs]                      # reference fbank values as input
    ref, rl = ofb.collate_frames([ofb.utterance_cmvn(r) for r in raw])
    m = ref.shape[1]
    feats = torch.zeros(B, m, 80)
    for i, r in enumerate(raw):
        feats[i, : r.shape[0]] = torch.from_numpy(r)
    feats, flens = feats.to(cuda), torch.from_numpy(rl).to(cuda)
    ms = torch.zeros(B, 2, 80, device=cuda)
    K.cmvn_stats(feats, flens, False, ms)
    o32 = torch.full((B, m, 80), 9.0, device=cuda)
    m_alloc = m + 4 + (m & 1)
    o16 = torch.full((B, m_alloc, 80), 9.0, dtype=torch.bfloat16, device=cuda)
    K.cmvn_apply(feats, ms, flens, False, o32, o16, op_row_offset=2)
    torch.cuda.synchronize()
    got = o32.cpu()
    assert torch.equal(got, torch.from_numpy(ref))
    assert torch.equal(o16[:, 2: 2 + m].cpu(), got.to(torch.bfloat16))
    assert o16[:, :2].abs().max().item() == 0 and o16[:, 2 + m:].abs().max().item() == 0


def test_fbank_cmvn_end_to_end(cuda):
    """Device fbank -> stats -> apply vs torchaudio + numpy CMVN (differences: fbank's 1e-6-level rounding only)."""
    from mm_s2ut_b200 import kernels as K
    from oracle import fbank as ofb

    wavs, (wav, lens) = _fbank_batch(0, 4, 2.0)
    wav, lens = wav.to(cuda), lens.to(cuda)
    B = len(wavs)
    m = 1 + (wav.shape[1] - 400) // 160
    feats = torch.zeros(B, m, 80, device=cuda)
    K.fbank(wav, lens, feats, K.fbank_tables(cuda))
    ms = torch.zeros(B, 2, 80, device=cuda)
    K.cmvn_stats(feats, lens, True, ms)
    o32 = torch.full((B, m, 80), 9.0, device=cuda)
    K.cmvn_apply(feats, ms, lens, True, o32, None)
    torch.cuda.synchronize()
    ref, rl = ofb.features_from_waveforms(wavs)
    assert ref.shape[1] == m
    # the raw-moment variance amplifies the 1e-6-level fbank differences in low-variance bins
    assert (o32.cpu() - torch.from_numpy(ref)).abs().max().item() < 3e-2


@pytest.mark.parametrize("T,lens", [(250, [250, 173, 1]), (125, [125, 80]), (300, [300, 257, 40]), (750, [750, 512]),
                                    (600, [600, 130, 128, 129, 1, 385]), (1030, [1030, 7, 512])])
def test_self_attention(cuda, T, lens):
    from mm_s2ut_b200 import kernels as K

    dt, H, hd = torch.bfloat16, 4, 64
    B, d = len(lens), 4 * 64
    g = torch.Generator().manual_seed(T)
    q = torch.randn(B, T, H, hd, generator=g) * 0.8
    k = torch.randn(B, T, H, hd, generator=g)
    v = torch.randn(B, T, H, hd, generator=g)
    qkv = torch.cat([q.reshape(B * T, d), k.reshape(B * T, d), v.reshape(B * T, d)], 1).to(cuda).to(dt).contiguous()
    sl = torch.tensor(lens, dtype=torch.int32, device=cuda)
    out = torch.zeros(B * T, d, dtype=dt, device=cuda)
    K.self_attention(qkv, sl, B, T, H, out)
    torch.cuda.synchronize()
    qf = qkv[:, :d].float().view(B, T, H, hd).permute(0, 2, 1, 3)
    kf = qkv[:, d: 2 * d].float().view(B, T, H, hd).permute(0, 2, 1, 3)
    vf = qkv[:, 2 * d:].float().view(B, T, H, hd).permute(0, 2, 1, 3)
    s = qf @ kf.transpose(-1, -2)
    mask = torch.arange(T, device=cuda)[None, :] >= sl[:, None]
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3).reshape(B * T, d)
    assert (out.float() - ref).abs().max().item() < 3e-2


def test_self_attention_many_items_per_cta(cuda):
    """BASELINE shape of the attention core: 64 utterances x 8 heads x 2 query tiles = 1024 items for 148 persistent CTAs
    (six or seven items each, so both softmax groups of a CTA are busy and hand the exponential sweep to each other in
    item order), ragged lengths, against plain fp32 PyTorch on the same 16-bit operands."""
    from mm_s2ut_b200 import kernels as K

    dt, B, T, H, hd = torch.bfloat16, 64, 250, 8, 64
    d = H * hd
    g = torch.Generator().manual_seed(64)
    lens = torch.randint(1, T + 1, (B,), generator=g)
    lens[0], lens[1] = T, 1
    qkv = torch.randn(B * T, 3 * d, generator=g)
    qkv[:, :d] *= 0.8
    qkv = qkv.to(cuda).to(dt).contiguous()
    sl = lens.to(torch.int32).to(cuda)
    out = torch.zeros(B * T, d, dtype=dt, device=cuda)
    for _ in range(3):      # the same bits every launch, whatever order the groups' turns were taken in
        prev = out.clone()
        K.self_attention(qkv, sl, B, T, H, out)
        torch.cuda.synchronize()
        assert _ == 0 or torch.equal(prev, out)
    qf, kf, vf = (qkv[:, i * d: (i + 1) * d].float().view(B, T, H, hd).permute(0, 2, 1, 3) for i in range(3))
    s = qf @ kf.transpose(-1, -2)
    mask = torch.arange(T, device=cuda)[None, :] >= sl[:, None]
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3).reshape(B * T, d)
    assert torch.isfinite(out).all()
    assert (out.float() - ref).abs().max().item() < 3e-2


def test_specaugment_fused_into_cmvn_matches_oracle(cuda):
    """CMVN + SpecAugment in one pass == oracle CMVN followed by the oracle (fairseq) SpecAugment, same draws."""
    import numpy as np

    from mm_s2ut_b200 import kernels as K
    from mm_s2ut_b200.data.specaugment import SpecAugmentTransform
    from oracle.fbank import utterance_cmvn
    from oracle.specaugment import spec_augment

    rng = np.random.RandomState(3)
    frames = [300, 211, 64]
    B, m = len(frames), max(frames)
    raw = rng.randn(B, m, 80).astype(np.float32) * 2 + 1
    sa = SpecAugmentTransform.from_policy("ld")
    tab = sa.draw_batch(frames, 80, np.random.RandomState(11))
    ref_rng = np.random.RandomState(11)
    feats = torch.from_numpy(raw).cuda()
    lens = torch.tensor(frames, dtype=torch.int64, device=cuda)
    stats = torch.empty(B, 2, 80, device=cuda)
    K.cmvn_stats(feats, lens, False, stats)
    out = torch.full((B, m, 80), 9.0, device=cuda)
    out_op = torch.full((B, m + 4, 80), 9.0, dtype=torch.bfloat16, device=cuda)
    K.cmvn_apply(feats, stats, lens, False, out, out_op, op_row_offset=2, spec_masks=torch.from_numpy(tab).cuda(),
                 n_fmask=sa.freq_mask_n, n_tmask=sa.time_mask_n, mask_value=sa.mask_value)
    for b, n in enumerate(frames):
        ref = spec_augment(utterance_cmvn(raw[b, :n]), freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100,
                           time_mask_p=1.0, mask_value=0.0, rng=ref_rng)
        got = out[b, :n].cpu().numpy()
        assert np.array_equal(got == 0.0, ref == 0.0)
        assert np.allclose(got, ref, atol=1e-5)
        assert (out[b, n:] == 0).all()
        assert torch.equal(out_op[b, 2:2 + n].float().cpu(), torch.from_numpy(got).bfloat16().float())
