"""Shared comparison helpers for the tests."""
import os

import numpy as np


def record(name: str, value, bound=None) -> None:
    """Append a measured parity number to $MM_PARITY_REPORT (profiles/rNN/parity.txt is produced this way)."""
    path = os.environ.get("MM_PARITY_REPORT")
    if path:
        with open(path, "a") as f:
            f.write(f"{name}: {value:.6g}" + (f"   (bound {bound:g})" if bound is not None else "") + "\n")


def fbank_errors(got: np.ndarray, ref: np.ndarray):
    """(err_main, err_all): max |got-ref| / max(|ref|, 1) over bins within 14 nats (~61 dB) of the frame maximum, and
    over ALL bins.  The north-star tolerance (1e-4 relative) is asserted on the former; bins more than 60 dB below the
    frame's strongest bin sit on the fp32 rounding floor of the reference's OWN rfft, where two faithful fp32
    implementations differ by a few 1e-4, so the all-bin figure gets the bound that is actually measured: 2e-4 for the
    CUDA kernel (worst of six utterances 1.0e-4, profiles/r01/parity.txt), 5e-4 for the numpy restatement."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    main = ref >= ref.max(axis=1, keepdims=True) - 14.0
    return float(rel[main].max()), float(rel.max())
