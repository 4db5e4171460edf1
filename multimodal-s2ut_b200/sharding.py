"""Length-bucketed utterance sharding across the GPUs of one box.

Forward / inference needs no collective: utterances are independent, so the work is partitioned exactly like
fairseq partitions it for the reference's training runs (reference launch flags
mm_s2ut/scripts/textless/1_train.sh:117-118 ``--max-tokens`` counted in fbank frames, ``--required-batch-size-multiple 1``):

  1. ``ordered_indices``  sort by frame count, descending, random tie-break (fairseq ``SpeechToTextDataset``)
  2. ``batch_by_size``    greedy token-budgeted batches: a batch costs (#utterances x longest utterance) frames
  3. ``shard_batches``    batch i goes to rank i % world (fairseq ``ShardedIterator``); every rank gets the same
                          number of batches (short ranks are padded with empty batches) so step counts agree.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def ordered_indices(n_frames: Sequence[int], seed: int = 1) -> np.ndarray:
    n_frames = np.asarray(n_frames)
    tie = np.random.RandomState(seed).permutation(len(n_frames))
    return np.lexsort((tie, -n_frames))


def batch_by_size(indices: Sequence[int], n_frames: Sequence[int], max_tokens: Optional[int] = None,
                  max_sentences: Optional[int] = None, required_batch_size_multiple: int = 1) -> List[List[int]]:
    """Greedy batching in index order; a sample that alone exceeds max_tokens still forms its own batch."""
    n_frames = np.asarray(n_frames)
    batches, cur, cur_max = [], [], 0
    for i in indices:
        ln = int(n_frames[i])
        new_max = max(cur_max, ln)
        over_tok = max_tokens is not None and cur and (len(cur) + 1) * new_max > max_tokens
        over_sent = max_sentences is not None and len(cur) >= max_sentences
        if over_tok or over_sent:
            keep = len(cur) - len(cur) % required_batch_size_multiple if len(cur) >= required_batch_size_multiple \
                else len(cur)
            batches.append(cur[:keep])
            cur = cur[keep:]
            cur_max = max((int(n_frames[j]) for j in cur), default=0)
            new_max = max(cur_max, ln)
        cur.append(int(i))
        cur_max = new_max
    if cur:
        batches.append(cur)
    return batches


def shard_batches(batches: List[List[int]], world_size: int, rank: int) -> List[List[int]]:
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    steps = (len(batches) + world_size - 1) // world_size
    mine = batches[rank::world_size]
    return mine + [[] for _ in range(steps - len(mine))]


def padding_fraction(batches: List[List[int]], n_frames: Sequence[int]) -> float:
    """Fraction of computed frames that are padding (what length bucketing minimises)."""
    n_frames = np.asarray(n_frames)
    real = sum(int(n_frames[b].sum()) for b in batches if b)
    comp = sum(len(b) * int(n_frames[b].max()) for b in batches if b)
    return 1.0 - real / max(comp, 1)
