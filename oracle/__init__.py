"""CPU oracle for the fbank -> fused-encoder hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package, and only as the checker or the timed CPU baseline.
The product package (``multimodal-s2ut_b200/``) never imports it.

What it restates (fp32, CPU, plain numpy / PyTorch), with the reference call site of each piece:

* ``fbank.py``    torchaudio ``compliance.kaldi.fbank`` as fairseq ``_get_torchaudio_fbank`` calls it
                  (reference: mm_s2ut/data/audio_utils.py:326-349), restated in numpy AND checked
                  against the real torchaudio installed in the image; fairseq ``UtteranceCMVN``
                  (reference: mm_s2ut/data/speech_to_speech_dataset.py:271-273); ``_collate_frames``
                  (reference: speech_to_speech_dataset.py:377-388).
* ``s2t.py``      fairseq ``Conv1dSubsampler``, sinusoidal positions, pre-LN ``TransformerEncoderLayer``
                  stack, final LayerNorm (reached through ``super().forward`` at
                  mm_s2ut/models/mm_s2s_transformer.py:464).
* ``fusion.py``   ``SelectiveAttention`` (mm_s2ut/models/fuse.py:35-117), ``MultimodalAttention``
                  (fuse.py:120-167), ``fuse_img_feat`` (mm_s2s_transformer.py:594-622), the
                  modality-dropout / sum glue (mm_s2s_transformer.py:496-530, :557-560).
* ``decoder.py``  fairseq ``TransformerUnitDecoder`` (consumer of the path, mm_s2s_transformer.py:693-696),
                  used on BOTH sides of the unit-argmax agreement check.

Third-party code that carries the arithmetic and is absent from /root/reference: fairseq
(facebookresearch/fairseq ``main``, un-pinned by the reference; >= Dec-2022) and torchaudio
(un-pinned; 2.11.0 in this image).

Pinning status: the reference has no tests or golden vectors ("parity unpinned" by the reference's
own tests).  The oracle is pinned instead against (1) real torchaudio 2.11 fbank outputs, (2) outputs
of the reference's own ``fuse.py`` imported in the build container (``oracle/make_golden.py`` ->
``tests/golden/*.npz``: the attention modules, their gradients, and the ``fuse_img_feat`` glue method), and (3) HF
``Speech2TextEncoder`` / ``Speech2TextDecoder`` -- independent ports of the same fairseq encoder and decoder -- with
copied weights; ``adam.py`` additionally against ``torch.optim.Adam`` in the eps -> 0 limit (the two differ only in where
eps enters).  fairseq itself could not be run (not installed, no network); the label-smoothed criterion, UtteranceCMVN
and SpecAugment are restated from its published source and stay unpinned.
"""
