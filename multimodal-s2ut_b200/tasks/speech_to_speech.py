"""Task ``multimodal_speech_to_speech`` (reference: mm_s2ut/tasks/speech_to_speech.py:45-123).

Only the boundary is mirrored: the registration name and the task-added flag
``--multimodal-translation-config-yaml`` that the encoder constructor reads.  Dataset construction (manifests,
image-feature stores, target units) stays with the reference's own data package, which is host-side I/O outside
the hot path; when fairseq is importable the task subclasses fairseq's ``SpeechToSpeechTask`` and otherwise this
module only exposes the flag definition for non-fairseq drivers."""
from __future__ import annotations

import argparse

TASK_NAME = "multimodal_speech_to_speech"


def add_multimodal_args(parser: argparse.ArgumentParser) -> None:
    parser.add_argument("--multimodal-translation-config-yaml", type=str, default=None,
                        help="YAML with the fusion keys (SA_image_dropout, use_selective_gate, image_feat_dim, ...)")
    parser.add_argument("--freezing-updates", type=int, default=None)


try:  # pragma: no cover - fairseq is not installed in the build image
    from fairseq.tasks import register_task
    from fairseq.tasks.speech_to_speech import SpeechToSpeechTask

    @register_task(TASK_NAME)
    class MultiModalSpeechToSpeechTask(SpeechToSpeechTask):
        @classmethod
        def add_args(cls, parser):
            super().add_args(parser)
            add_multimodal_args(parser)
except Exception:
    MultiModalSpeechToSpeechTask = None
