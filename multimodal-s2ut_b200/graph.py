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
                 wav_dtype: torch.dtype = torch.float32, stores: Optional[list] = None,
                 img_dtype: torch.dtype = torch.float32):
        """stores: optional list (one entry per image type) of ImageFeatureStore or None.  For a store the static
        input is a [batch] int64 index vector instead of a [batch, Tk, Dk] feature tensor.  img_dtype: float32 (the
        reference's collater output) or float16 / bfloat16 (features shipped in 16 bit: half the H2D bytes)."""
        if enc.training:
            raise RuntimeError("graph capture is for eval-mode forwards (modality dropout draws are per batch)")
        self.enc = enc
        dev = next(enc.parameters()).device
        self.device = dev
        self.wav = torch.zeros(batch, n_samples, dtype=wav_dtype, device=dev)   # float32 (x 2**15) or int16 PCM
        self.lens = torch.full((batch,), n_samples, dtype=torch.int64, device=dev)
        self.stores = list(stores) if stores is not None else [None for _ in img_shapes]
        self.imgs = [torch.zeros(batch, *s, dtype=img_dtype, device=dev) if st is None else
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


class GraphedTrainStep:
    """CUDA-graph replay of the training step (``training.TrainEngine``) for a fixed batch shape.

    Two graphs hold forward (activations kept) + backward, one per outcome of the per-batch modality-dropout draw
    (images kept / every image tensor zeroed, reference mm_s2s_transformer.py:496-505); a third holds gradient
    clipping + fairseq Adam + the operand refresh.  The gradient all-reduce runs between them (NCCL, outside the
    graphs).  Per-step hyper-parameters (Adam step size, weight decay x lr, gradient scale, clip norm) are written to
    the device before the optimizer graph is replayed, so learning-rate schedules and fairseq's
    ``world_size / sample_size`` gradient normalisation work under replay.
    """

    def __init__(self, enc, batch: int, n_samples: int, img_shape: tuple, wav_dtype: torch.dtype = torch.float32,
                 betas=(0.9, 0.98), eps: float = 1e-8, overlap_reduce: bool = False, specaugment=None):
        """specaugment: optional ``data.specaugment.SpecAugmentTransform``; its per-utterance masks are drawn on the host
        before every replay (``enc.modality_rng``, the reference's numpy calls) and copied into a static device table."""
        self.enc = enc
        self.eng = enc.train_engine()
        dev = self.eng.device
        self.device = dev
        self.betas, self.eps = betas, eps
        self.overlap_reduce = overlap_reduce
        self.wav = torch.zeros(batch, n_samples, dtype=wav_dtype, device=dev)
        self.lens = torch.full((batch,), n_samples, dtype=torch.int64, device=dev)
        self.img = torch.zeros(batch, *img_shape, dtype=torch.float32, device=dev)
        self.grad_out: Optional[torch.Tensor] = None
        # element-wise dropout under replay: the masks are keyed on base seed + this device scalar, advanced per step
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.base_seed = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF
        self.specaugment = specaugment
        self.spec_tab = None
        if specaugment is not None:
            n_masks = specaugment.freq_mask_n + specaugment.time_mask_n
            self.spec_tab = torch.zeros(batch, 2 * n_masks, dtype=torch.int32, device=dev)
            self._spec_frames = [0 if n_samples < 400 else 1 + (n_samples - 400) // 160] * batch
        self.out = {}
        self.graphs = {}
        self.opt_graph: Optional[torch.cuda.CUDAGraph] = None
        # pinned staging ring for the per-step hyper-parameters: a slot is rewritten only after its H2D copy has run
        self._hyper_host = [torch.zeros(4, dtype=torch.float32).pin_memory() for _ in range(8)]
        self._hyper_done = [None] * 8
        self._hyper_i = 0

    def _specaug(self):
        sa = self.specaugment
        return None if sa is None else (self.spec_tab, sa.freq_mask_n, sa.time_mask_n, sa.mask_value)

    def _fwd_bwd(self, drop_image: bool):
        out = self.eng.forward_train(self.wav, self.lens, [self.img], [None], drop_image=drop_image,
                                     dropout_seed=self.base_seed, dropout_seed_dev=self.seed_dev, specaug=self._specaug())
        if self.grad_out is None:
            self.grad_out = torch.zeros_like(out["encoder_out"][0])
        # with several ranks the bucketed gradient all-reduce is part of the captured graph (side stream, forked from
        # and joined to the capture stream), overlapping the rest of the backward pass
        self.eng.backward(self.grad_out, overlap_reduce=self.overlap_reduce)
        return out

    def capture(self) -> None:
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):          # eager pass: allocates every workspace, sets kernel attributes
            for drop in (False, True):
                self._fwd_bwd(drop)
            for nc in self._hyper_targets():
                nc[2:6] = torch.tensor([0.0, 0.0, 1.0, 0.0], device=self.device)   # a no-op optimizer step
            self._optimizer()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self._reset_moments()
        pool = None
        for drop in (False, True):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                self.out[drop] = self._fwd_bwd(drop)
            pool = g.pool()
            self.graphs[drop] = g
        self.opt_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.opt_graph, pool=pool):
            self._optimizer()

    # hooks the encoder + decoder step overrides
    def _optimizer(self) -> None:
        self.eng.adam_step_device_hyper(self.betas, self.eps)

    def _hyper_targets(self):
        return [self.eng.norm_coef]

    def _reset_moments(self) -> None:
        self.eng.exp_avg.zero_()
        self.eng.exp_avg_sq.zero_()

    def forward_backward(self, drop_image: bool = False):
        """Replays forward + backward on the static inputs (``wav``, ``lens``, ``img``, ``grad_out``)."""
        if self.opt_graph is None:
            self.capture()
        self.seed_dev.add_(0x632BE5AB)        # fresh dropout masks for this step (stream-ordered before the replay)
        if self.specaugment is not None:      # fresh SpecAugment masks: host draws -> the static device table
            tab = self.specaugment.draw_batch(self._spec_frames, 80, self.enc.modality_rng)
            self.spec_tab.copy_(torch.from_numpy(tab))
        self.graphs[bool(drop_image)].replay()
        return self.out[bool(drop_image)]

    def optimizer_step(self, lr: float, weight_decay: float = 0.0, clip_norm: float = 0.0,
                       grad_scale: float = 1.0) -> None:
        vals = self.eng.hyper_values(lr, self.betas, weight_decay, clip_norm, grad_scale)
        i = self._hyper_i
        self._hyper_i = (i + 1) % len(self._hyper_host)
        if self._hyper_done[i] is not None:
            self._hyper_done[i].synchronize()
        self._hyper_host[i].copy_(torch.tensor(vals, dtype=torch.float32))
        for k, nc in enumerate(self._hyper_targets()):
            nc[2:6].copy_(self._hyper_host[i], non_blocking=True)
            if k > 0:
                nc[5].zero_()     # only the first (encoder) engine clips: it holds the joint norm
        ev = torch.cuda.Event()
        ev.record()
        self._hyper_done[i] = ev
        self.opt_graph.replay()

    def step(self, lr: float, drop_image: bool = False, weight_decay: float = 0.0, clip_norm: float = 0.0):
        """forward + backward -> gradient all-reduce -> optimizer, on the static inputs."""
        out = self.forward_backward(drop_image)
        import torch.distributed as dist

        if self.overlap_reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.eng._reduced = True      # the captured graph holds the collectives (python state is not replayed)
        ws = self.eng.all_reduce_grads()
        self.optimizer_step(lr, weight_decay, clip_norm, grad_scale=1.0 / ws)
        return out


class GraphedModelTrainStep(GraphedTrainStep):
    """The complete BASELINE configs[2] step under CUDA-graph replay: encoder (``TrainEngine``) + S2UT unit decoder +
    label-smoothed cross entropy (``decoder_training.UnitDecoderTrainEngine``), chained by ``d loss / d encoder_out``
    with no autograd:  waveform -> fused states -> logits -> loss -> decoder backward -> encoder backward ->
    (gradient all-reduce) -> joint-norm clipping -> Adam on both engines.  Static inputs: ``wav``, ``lens``, ``img``,
    ``prev_tokens``, ``target``."""

    def __init__(self, enc, dec_engine, batch: int, n_samples: int, img_shape: tuple, tgt_len: int,
                 label_smoothing: float = 0.2, **kw):
        super().__init__(enc, batch, n_samples, img_shape, **kw)
        self.dec = dec_engine
        self.label_smoothing = label_smoothing
        self.prev_tokens = torch.full((batch, tgt_len), 4, dtype=torch.int64, device=self.device)
        self.target = torch.full((batch, tgt_len), 4, dtype=torch.int64, device=self.device)
        self.loss = {}
        self.grad_out = False      # unused: the decoder supplies d loss / d encoder_out

    def _fwd_bwd(self, drop_image: bool):
        import torch.distributed as dist

        out = self.eng.forward_train(self.wav, self.lens, [self.img], [None], drop_image=drop_image,
                                     dropout_seed=self.base_seed, dropout_seed_dev=self.seed_dev, specaug=self._specaug())
        self.dec.forward_train(self.prev_tokens, out["encoder_out"][0], out["encoder_padding_mask"][0],
                               dropout_seed=self.base_seed, dropout_seed_dev=self.seed_dev)
        loss, nll, d_enc = self.dec.loss_backward(self.target, self.label_smoothing)
        multi = self.overlap_reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if multi:       # the decoder's gradients travel while the encoder's backward runs
            self.eng._reduce_works = []
            ev = torch.cuda.Event()
            ev.record()
            if self.eng._comm_stream is None:
                self.eng._comm_stream = torch.cuda.Stream(device=self.device)
            from .peer import peer_group

            grp = peer_group(self.dec.flat_g)
            with torch.cuda.stream(self.eng._comm_stream):
                self.eng._comm_stream.wait_event(ev)
                if grp is not None:
                    grp.all_reduce()
                else:
                    self.eng._reduce_works.append(dist.all_reduce(self.dec.flat_g, op=dist.ReduceOp.SUM, async_op=True))
        self.eng.backward(d_enc, overlap_reduce=self.overlap_reduce)
        self.loss[bool(drop_image)] = (loss, nll)
        return out

    def _optimizer(self) -> None:
        self.dec.grad_norm(dev_hyper=True)
        self.eng.adam_step_device_hyper(self.betas, self.eps, extra_norm=self.dec.norm_coef)
        self.dec.adam_apply(self.eng.norm_coef, betas=self.betas, eps=self.eps, step=0)

    def _hyper_targets(self):
        return [self.eng.norm_coef, self.dec.norm_coef]

    def _reset_moments(self) -> None:
        super()._reset_moments()
        self.dec.exp_avg.zero_()
        self.dec.exp_avg_sq.zero_()

    def step(self, lr: float, drop_image: bool = False, weight_decay: float = 0.0, clip_norm: float = 0.0):
        import torch.distributed as dist

        out = self.forward_backward(drop_image)
        ws = 1
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            ws = dist.get_world_size()
            if not self.overlap_reduce:
                from .training import all_reduce_flat

                self.eng.all_reduce_grads()
                all_reduce_flat(self.dec.flat_g)
        self.optimizer_step(lr, weight_decay, clip_norm, grad_scale=1.0 / ws)
        return out, self.loss[bool(drop_image)]
