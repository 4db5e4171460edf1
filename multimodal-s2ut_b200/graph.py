"""CUDA-graph replay of the encoder forward for fixed batch shapes.

The forward is ~150 small-to-medium kernel launches; issuing them from Python costs more host time than the
GPU needs to run them.  ``GraphedEncoder`` captures one eager forward (all launches go through the C ABI on
the capturing stream; workspaces are the engine's cached buffers, so no allocation happens inside the
graph except the result tensors, which live in the graph's private pool) and replays it with one
``cudaGraphLaunch``.  Inputs are copied into static device buffers (from pinned host memory when the caller
has host tensors); outputs are the static tensors of the captured run.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .feature_store import StoredImages


class GraphedEncoder:
    def __init__(self, enc, batch: int, n_samples: int, img_shapes: List[tuple], warmup: int = 2,
                 wav_dtype: torch.dtype = torch.float32, stores: Optional[list] = None):
        """stores: optional list (one entry per image type) of ImageFeatureStore or None.  For a store the static
        input is a [batch] int64 index vector instead of a [batch, Tk, Dk] fp32 feature tensor."""
        if enc.training:
            raise RuntimeError("graph capture is for eval-mode forwards (modality dropout draws are per batch)")
        self.enc = enc
        dev = next(enc.parameters()).device
        self.device = dev
        self.wav = torch.zeros(batch, n_samples, dtype=wav_dtype, device=dev)   # float32 (x 2**15) or int16 PCM
        self.lens = torch.full((batch,), n_samples, dtype=torch.int64, device=dev)
        self.stores = list(stores) if stores is not None else [None for _ in img_shapes]
        self.imgs = [torch.zeros(batch, *s, dtype=torch.float32, device=dev) if st is None else
                     StoredImages(st, torch.zeros(batch, dtype=torch.int64, device=dev))
                     for s, st in zip(img_shapes, self.stores)]
        self.masks: List[Optional[torch.Tensor]] = [None for _ in img_shapes]
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.out = None
        self._warmup = warmup

    def _run(self):
        return self.enc(self.wav, self.lens, None, None, None, imgs_list=list(self.imgs),
                        img_masks_list=list(self.masks))

    def capture(self) -> None:
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(self._warmup):   # allocates workspaces, sets kernel attributes
                self._run()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._run()

    def load_inputs(self, wav: torch.Tensor, lens: torch.Tensor, imgs: List[torch.Tensor]) -> None:
        """Async copies into the static buffers on the current stream (H2D when the sources are pinned host)."""
        self.wav.copy_(wav, non_blocking=True)
        self.lens.copy_(lens, non_blocking=True)
        for dst, src in zip(self.imgs, imgs):   # feature tensors, or index vectors for store-backed image types
            (dst.index if isinstance(dst, StoredImages) else dst).copy_(src, non_blocking=True)

    def replay(self):
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.out

    def __call__(self, wav, lens, imgs):
        self.load_inputs(wav, lens, imgs)
        return self.replay()
