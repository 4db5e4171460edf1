"""Oracle (test infrastructure): fairseq ``SpecAugmentTransform.__call__`` restated in numpy.

Reference call site: mm_s2ut/data/speech_to_speech_dataset.py:271-273 (``self.feature_transforms(source)`` with the
``_train: [utterance_cmvn, specaugment]`` transform list of the data config).  fairseq
(``fairseq/data/audio/feature_transforms/specaugment.py``) is un-vendored and absent here: this restates its
published algorithm -- **parity unpinned** by reference tests.  ``mask_value=None`` means the spectrogram mean.
"""
from __future__ import annotations

import math

import numpy as np


def spec_augment(spectrogram: np.ndarray, *, freq_mask_n=0, freq_mask_f=0, time_mask_n=0, time_mask_t=0, time_mask_p=0.0,
                 mask_value=0.0, rng=np.random) -> np.ndarray:
    assert spectrogram.ndim == 2
    distorted = spectrogram.copy()
    num_frames, num_freqs = spectrogram.shape
    if mask_value is None:
        mask_value = spectrogram.mean()
    if num_frames == 0:
        return spectrogram
    if num_freqs < freq_mask_f:
        return spectrogram
    for _ in range(freq_mask_n):
        f = rng.randint(0, freq_mask_f)
        f0 = rng.randint(0, num_freqs - f)
        if f != 0:
            distorted[:, f0:f0 + f] = mask_value
    max_time_mask_t = min(time_mask_t, math.floor(num_frames * time_mask_p))
    if max_time_mask_t < 1:
        return distorted
    for _ in range(time_mask_n):
        t = rng.randint(0, max_time_mask_t)
        t0 = rng.randint(0, num_frames - t)
        if t != 0:
            distorted[t0:t0 + t, :] = mask_value
    return distorted
