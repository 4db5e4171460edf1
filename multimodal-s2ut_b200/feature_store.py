"""Device-resident image-feature store.

The reference keeps the whole ``[N, 577, 768]`` fp32 feature tensor of a split in host RAM
(``ImageDataset``: ``torch.load(feat_path)``, ``__getitem__`` returns ``img_feat[idx]``,
mm_s2ut/data/speech_to_speech_dataset.py:36-68) and ships 1.77 MB per utterance over PCIe every step; with the
kernels of this repo that copy (113 MB per 64-utterance batch) is what bounds the end-to-end rate.  Multi30k's
29 000 x 577 x 768 features are 25.7 GB in 16 bit: they fit in one B200's HBM several times over.

``ImageFeatureStore`` holds the features on the GPU in 16 bit (fp16 by default: ViT features are O(1), and fp16's
11-bit significand keeps the rounding of the inputs 8x below that of the bf16 GEMM operands they become).  A batch is
then just an index vector: ``store.batch(indices)`` returns a handle that the encoder accepts in ``imgs_list`` in place
of a ``[B, Tk, Dk]`` tensor; the image pre-norm kernel gathers the rows straight from the store
(``mm_layernorm_gather``), so the batch is never materialised and the store is read once per step, in 16 bit.
"""
from __future__ import annotations

from typing import Optional

import torch


class StoredImages:
    """A batch of image features that lives in an ImageFeatureStore: (store, int64 device index vector)."""

    def __init__(self, store: "ImageFeatureStore", index: torch.Tensor):
        assert index.dtype == torch.int64 and index.dim() == 1
        self.store, self.index = store, index

    @property
    def shape(self):
        return (self.index.numel(), self.store.tokens, self.store.dim)


class ImageFeatureStore:
    def __init__(self, feats: torch.Tensor, device, dtype: torch.dtype = torch.float16, chunk: int = 1024):
        """feats [N, Tk, Dk] (any float dtype, host or device) -> 16-bit copy on `device` (converted chunk by chunk so
        the host tensor never needs a full-size temporary)."""
        if feats.dim() != 3:
            raise ValueError("image features must be [N, tokens, dim]")
        if dtype not in (torch.float16, torch.bfloat16):
            raise TypeError("the store holds float16 or bfloat16")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ImageFeatureStore lives on a CUDA device (no CPU fallback exists)")
        n, self.tokens, self.dim = feats.shape
        self.data = torch.empty(n, self.tokens, self.dim, dtype=dtype, device=self.device)
        for i in range(0, n, chunk):
            self.data[i:i + chunk].copy_(feats[i:i + chunk].to(self.device, non_blocking=True))

    def __len__(self) -> int:
        return self.data.shape[0]

    def batch(self, indices, out: Optional[torch.Tensor] = None) -> StoredImages:
        """indices: int sequence / tensor of dataset positions (the sample ids the reference's collater gathers)."""
        idx = torch.as_tensor(indices, dtype=torch.int64)
        if idx.device.type == "cpu" and idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= len(self)):
            raise IndexError(f"image index out of range for a store of {len(self)} samples")   # the gather trusts them
        if out is not None:
            out.copy_(idx, non_blocking=True)
            idx = out
        else:
            idx = idx.to(self.device, non_blocking=True)
        return StoredImages(self, idx)
