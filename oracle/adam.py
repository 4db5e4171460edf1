"""Oracle (test infrastructure): fairseq's Adam step and gradient clipping, numpy fp32/fp64.

The reference trains with ``--optimizer adam --adam-betas '(0.9,0.98)' --clip-norm 10.0`` (scripts/textless/1_train.sh
:111-113 region; fairseq's own ``fairseq/optim/adam.py`` ``Adam.step`` and ``fairseq/utils.py`` ``clip_grad_norm_`` --
fairseq is un-vendored and absent here, so this restates its published algorithm: **parity unpinned** by reference
tests; checked against ``torch.optim.Adam`` in the eps -> 0 limit, tests/test_oracle_golden.py).  fairseq's step
differs from ``torch.optim.Adam`` only in where eps enters:

    exp_avg    = b1 * exp_avg    + (1 - b1) * g
    exp_avg_sq = b2 * exp_avg_sq + (1 - b2) * g * g
    denom      = sqrt(exp_avg_sq) + eps
    step_size  = lr * sqrt(1 - b2^t) / (1 - b1^t)
    p          = p - weight_decay * lr * p          (if weight_decay != 0)
    p          = p - step_size * exp_avg / denom
"""
from __future__ import annotations

import numpy as np


def clip_coef(grad: np.ndarray, grad_scale: float, max_norm: float):
    """(norm, multiplier): fairseq multiply_grads(grad_scale) then clip_grad_norm_(max_norm)."""
    norm = float(np.sqrt(np.sum(grad.astype(np.float64) ** 2))) * grad_scale
    coef = grad_scale
    if max_norm > 0:
        coef *= min(1.0, max_norm / (norm + 1e-6))
    return norm, coef


def adam_step(p, g, m, v, *, lr, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.0, step=1, grad_mul=1.0):
    b1, b2 = betas
    g = (g * np.float32(grad_mul)).astype(np.float32)
    m = (np.float32(b1) * m + np.float32(1 - b1) * g).astype(np.float32)
    v = (np.float32(b2) * v + np.float32(1 - b2) * g * g).astype(np.float32)
    step_size = np.float32(lr * np.sqrt(1 - b2 ** step) / (1 - b1 ** step))
    p = p.astype(np.float32)
    if weight_decay != 0:
        p = p - np.float32(weight_decay * lr) * p
    p = (p - step_size * m / (np.sqrt(v) + np.float32(eps))).astype(np.float32)
    return p, m, v
