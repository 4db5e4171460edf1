"""Seeded synthetic inputs for the fbank -> fused-encoder hot path.

The recipe is SURVEY.md §8(d): per utterance ``u`` of config ``c`` a generator
seeded with ``1234 + 1000*c + u``; waveform = clip(0.1*N(0,1) + sum of 5
sinusoids, -1, 1) scaled by 2**15 (what the reference's ``get_waveform`` does at
mm_s2ut/data/audio_utils.py:289-290 before calling fbank); image features are
N(0,1) ``[B, 577, 768]`` (ViT-base patch tokens, the shape
mm_s2ut/scripts/extract_feature/get_img_feat_vit.py:69-107 writes).

Everything here is host-side numpy/torch on the CPU; it is shared by the tests,
``bench.py`` and the oracle so that every arm sees identical inputs.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np
import torch

SAMPLE_RATE = 16000


def synth_waveform(config_idx: int, utt_idx: int, dur_s: float, ragged: bool = True,
                   all_zero: bool = False) -> np.ndarray:
    """One utterance, float32, already multiplied by 2**15 (Kaldi-style int16 range)."""
    g = torch.Generator().manual_seed(1234 + 1000 * config_idx + utt_idx)
    frac = 0.6 + 0.4 * torch.rand(1, generator=g).item() if ragged else 1.0
    n = int(round(SAMPLE_RATE * dur_s * frac))
    if all_zero:
        return np.zeros(n, dtype=np.float32)
    t = torch.arange(n, dtype=torch.float64) / SAMPLE_RATE
    wav = 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    for _ in range(5):
        f = 80.0 + (4000.0 - 80.0) * torch.rand(1, generator=g).item()
        a = 0.02 + (0.2 - 0.02) * torch.rand(1, generator=g).item()
        ph = 2.0 * math.pi * torch.rand(1, generator=g).item()
        wav = wav + a * torch.sin(2.0 * math.pi * f * t + ph)
    wav = wav.clamp_(-1.0, 1.0) * 32768.0
    return wav.to(torch.float32).numpy()


def synth_batch(config_idx: int, n_utts: int, dur_s: float, ragged: bool = True,
                zero_utt: Optional[int] = None) -> Tuple[List[np.ndarray], np.ndarray]:
    """List of waveforms plus their sample counts."""
    wavs = [synth_waveform(config_idx, u, dur_s, ragged, all_zero=(zero_utt == u)) for u in range(n_utts)]
    return wavs, np.array([len(w) for w in wavs], dtype=np.int64)


def pad_waveforms(wavs: List[np.ndarray]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Zero-pad to ``[B, Nmax]`` float32 + int64 lengths (the raw-waveform input contract)."""
    n = max(len(w) for w in wavs)
    out = torch.zeros(len(wavs), n, dtype=torch.float32)
    for i, w in enumerate(wavs):
        out[i, : len(w)] = torch.from_numpy(w)
    return out, torch.tensor([len(w) for w in wavs], dtype=torch.int64)


def synth_images(config_idx: int, n_utts: int, n_tokens: int = 577, dim: int = 768) -> torch.Tensor:
    g = torch.Generator().manual_seed(987654 + 1000 * config_idx)
    return torch.randn(n_utts, n_tokens, dim, generator=g, dtype=torch.float32)


def num_frames(n_samples: int) -> int:
    """snip_edges=True frame count (torchaudio/compliance/kaldi.py:67)."""
    return 0 if n_samples < 400 else 1 + (n_samples - 400) // 160


def subsampled_len(n: int, n_layers: int = 2) -> int:
    """fairseq Conv1dSubsampler.get_out_seq_lens_tensor: floor((L-1)/2 + 1) per layer."""
    for _ in range(n_layers):
        n = (n - 1) // 2 + 1
    return n
