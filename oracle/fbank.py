"""Oracle (test infrastructure): Kaldi-compatible fbank, utterance CMVN, frame collation.

Follows torchaudio/compliance/kaldi.py (2.11.0) with the arguments fairseq's
``_get_torchaudio_fbank`` passes: ``fbank(waveform, num_mel_bins=80, sample_frequency=16000)``,
everything else default (reference call site: mm_s2ut/data/audio_utils.py:341-343).
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

EPS = np.float32(1.1920928955078125e-07)  # torch.finfo(float32).eps, kaldi.py:_get_epsilon


def mel_scale(f):
    return 1127.0 * np.log(1.0 + f / 700.0)


def povey_window(n: int = 400) -> np.ndarray:
    """hann(n, periodic=False) ** 0.85 in fp32 (kaldi.py:96-100)."""
    k = np.arange(n, dtype=np.float64)
    hann = (0.5 - 0.5 * np.cos(2.0 * math.pi * k / (n - 1))).astype(np.float32)
    return np.power(hann, np.float32(0.85)).astype(np.float32)


def mel_banks(num_bins: int = 80, n_fft: int = 512, sample_rate: float = 16000.0,
              low_freq: float = 20.0, high_freq: float = 0.0) -> np.ndarray:
    """``get_mel_banks`` (kaldi.py:400-502), vtln_warp == 1.0; returns [num_bins, n_fft//2] fp32.

    torchaudio evaluates the slopes on fp32 tensors; we do the same so the bank is bit-close.
    """
    import torch  # fp32 tensor arithmetic identical to torchaudio's

    num_fft_bins = n_fft / 2
    nyquist = 0.5 * sample_rate
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_rate / n_fft
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + fft_bin_width * torch.arange(num_fft_bins) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    return bins.numpy().astype(np.float32)


def kaldi_fbank_np(wave: np.ndarray, num_mel_bins: int = 80) -> np.ndarray:
    """Numpy restatement. ``wave``: float32 [n], already scaled by 2**15. Returns [m, 80] fp32."""
    wave = np.asarray(wave, dtype=np.float32)
    n = wave.shape[0]
    win, shift, nfft = 400, 160, 512
    if n < win:
        return np.zeros((0, num_mel_bins), dtype=np.float32)
    m = 1 + (n - win) // shift                                    # snip_edges=True, kaldi.py:67
    idx = np.arange(m)[:, None] * shift + np.arange(win)[None, :]
    fr = wave[idx]                                                # [m, 400]
    fr = fr - fr.mean(axis=1, keepdims=True, dtype=np.float32)    # remove_dc_offset, kaldi.py:185-186
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)        # replicate pad, kaldi.py:195-198
    fr = fr - np.float32(0.97) * prev
    fr = fr * povey_window(win)[None, :]                          # kaldi.py:201-204
    pad = np.zeros((m, nfft), dtype=np.float32)
    pad[:, :win] = fr
    import scipy.fft                                              # pocketfft in single precision, like torch's CPU rfft

    spec = scipy.fft.rfft(pad, axis=1)                            # complex64: the reference's rfft is fp32
    power = (np.abs(spec) ** np.float32(2.0)).astype(np.float32)  # [m, 257]  (kaldi.py:616-618: abs().pow(2))
    bank = np.pad(mel_banks(num_mel_bins, nfft), ((0, 0), (0, 1)))  # Nyquist column zero, kaldi.py:627
    mel = power @ bank.T
    return np.log(np.maximum(mel, EPS)).astype(np.float32)        # kaldi.py:633


def kaldi_fbank_ta(wave: np.ndarray, num_mel_bins: int = 80) -> np.ndarray:
    """The real thing: exactly what fairseq ``_get_torchaudio_fbank`` executes."""
    import torch
    import torchaudio.compliance.kaldi as ta_kaldi

    w = torch.from_numpy(np.asarray(wave, dtype=np.float32)).unsqueeze(0)
    return ta_kaldi.fbank(w, num_mel_bins=num_mel_bins, sample_frequency=16000).numpy()


def utterance_cmvn(x: np.ndarray, norm_means: bool = True, norm_vars: bool = True) -> np.ndarray:
    """fairseq ``UtteranceCMVN.__call__``: raw-moment population variance with a 1e-10 floor."""
    x = np.asarray(x, dtype=np.float32)
    mean = x.mean(axis=0)
    square_sums = (x ** 2).sum(axis=0)
    if norm_means:
        x = np.subtract(x, mean)
    if norm_vars:
        var = square_sums / x.shape[0] - mean ** 2
        std = np.sqrt(np.maximum(var, 1e-10))
        x = np.divide(x, std)
    return x.astype(np.float32)


def collate_frames(frames: List[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """fairseq ``_collate_frames``: zero-pad AFTER CMVN to the batch max length."""
    lens = np.array([f.shape[0] for f in frames], dtype=np.int64)
    out = np.zeros((len(frames), int(lens.max()), frames[0].shape[1]), dtype=np.float32)
    for i, f in enumerate(frames):
        out[i, : f.shape[0]] = f
    return out, lens


def features_from_waveforms(wavs: List[np.ndarray], use_torchaudio: bool = True):
    """Per-utterance fbank + CMVN (sequential, like the reference ``__getitem__``), then collate."""
    fb = kaldi_fbank_ta if use_torchaudio else kaldi_fbank_np
    feats = [utterance_cmvn(fb(w)) for w in wavs]
    return collate_frames(feats)
