"""Host side of the on-device SpecAugment.

The reference applies fairseq's ``SpecAugmentTransform`` per utterance on the CPU, after ``utterance_cmvn``, in
training only (data config ``transforms: {"_train": [utterance_cmvn, specaugment]}``, called at
mm_s2ut/data/speech_to_speech_dataset.py:271-273).  With fbank + CMVN on the GPU the augmentation moves there too:
this class keeps the transform's configuration surface (``from_config_dict`` keys, the named policies of fairseq's
``gen_config_yaml``) and DRAWS the masks with the same ``np.random.randint`` calls in the same order; the masking
itself is fused into the CMVN kernel (``mm_cmvn_apply_specaug``).  Time warping (``time_warp_W > 0``, needs OpenCV in
fairseq; 0 in every named policy but "ld"/"sm"/"ss" use 0 as well) is not supported.

``mask_value``: fairseq's config path leaves it ``None`` = the utterance's mean after CMVN, which is 0 up to fp32
rounding (~1e-7); 0.0 is written.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

POLICIES = {   # fairseq examples/speech_to_text/data_utils.py: set_specaugment_{lb,ld,sm,ss}_policy
    "lb": dict(time_warp_W=0, freq_mask_N=1, freq_mask_F=27, time_mask_N=1, time_mask_T=100, time_mask_p=1.0),
    "ld": dict(time_warp_W=0, freq_mask_N=2, freq_mask_F=27, time_mask_N=2, time_mask_T=100, time_mask_p=1.0),
    "sm": dict(time_warp_W=0, freq_mask_N=2, freq_mask_F=15, time_mask_N=2, time_mask_T=70, time_mask_p=0.2),
    "ss": dict(time_warp_W=0, freq_mask_N=2, freq_mask_F=27, time_mask_N=2, time_mask_T=70, time_mask_p=0.2),
}


class SpecAugmentTransform:
    def __init__(self, time_warp_w: int = 0, freq_mask_n: int = 0, freq_mask_f: int = 0, time_mask_n: int = 0,
                 time_mask_t: int = 0, time_mask_p: float = 0.0, mask_value: Optional[float] = 0.0):
        if time_warp_w > 0:
            raise NotImplementedError("time warping is not supported on the device path (0 in fairseq's policies)")
        if freq_mask_n > 0 and freq_mask_f <= 0:
            raise ValueError(f"freq_mask_F ({freq_mask_f}) must be larger than 0 when doing freq masking")
        if time_mask_n > 0 and time_mask_t <= 0:
            raise ValueError(f"time_mask_T ({time_mask_t}) must be larger than 0 when doing time masking")
        self.freq_mask_n, self.freq_mask_f = freq_mask_n, freq_mask_f
        self.time_mask_n, self.time_mask_t, self.time_mask_p = time_mask_n, time_mask_t, time_mask_p
        self.mask_value = 0.0 if mask_value is None else float(mask_value)

    @classmethod
    def from_config_dict(cls, config: Optional[Dict] = None):
        c = {} if config is None else config
        return cls(c.get("time_warp_W", 0), c.get("freq_mask_N", 0), c.get("freq_mask_F", 0), c.get("time_mask_N", 0),
                   c.get("time_mask_T", 0), c.get("time_mask_p", 0.0), c.get("mask_value", None))

    @classmethod
    def from_policy(cls, name: str):
        return cls.from_config_dict(POLICIES[name])

    def draw(self, num_frames: int, num_freqs: int = 80, rng=np.random) -> Tuple[List[Tuple[int, int]], List[Tuple[int, int]]]:
        """The (f0, f) and (t0, t) masks of ONE utterance, drawn exactly like ``SpecAugmentTransform.__call__``
        (same calls, same order, same early exits); width 0 = no mask."""
        fm = [(0, 0)] * self.freq_mask_n
        tm = [(0, 0)] * self.time_mask_n
        if num_frames == 0 or num_freqs < self.freq_mask_f:
            return fm, tm
        for i in range(self.freq_mask_n):
            f = int(rng.randint(0, self.freq_mask_f))
            f0 = int(rng.randint(0, num_freqs - f))
            fm[i] = (f0, f)
        max_t = min(self.time_mask_t, math.floor(num_frames * self.time_mask_p))
        if max_t < 1:
            return fm, tm
        for i in range(self.time_mask_n):
            t = int(rng.randint(0, max_t))
            t0 = int(rng.randint(0, num_frames - t))
            tm[i] = (t0, t)
        return fm, tm

    def draw_batch(self, frames: List[int], num_freqs: int = 80, rng=np.random) -> np.ndarray:
        """int32 [B, 2 (freq_mask_N + time_mask_N)] mask table of a batch (utterances in batch order)."""
        out = np.zeros((len(frames), 2 * (self.freq_mask_n + self.time_mask_n)), dtype=np.int32)
        for b, m in enumerate(frames):
            fm, tm = self.draw(int(m), num_freqs, rng)
            out[b] = [v for pair in fm + tm for v in pair]
        return out
