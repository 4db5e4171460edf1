"""Task ``multimodal_speech_to_speech`` (reference: mm_s2ut/tasks/speech_to_speech.py:45-123).

Only the boundary is mirrored: the registration name, the task-added flag
``--multimodal-translation-config-yaml`` that the encoder constructor reads, and ``load_dataset``, which makes the
reference's own call into the reference's own data package (manifests, image-feature tensors, target units and the
collater are host-side I/O outside the hot path; the package must be importable).  When fairseq is importable the task
subclasses fairseq's ``SpeechToSpeechTask``; otherwise this module only exposes the flag definition for non-fairseq
drivers."""
from __future__ import annotations

import argparse

TASK_NAME = "multimodal_speech_to_speech"


def add_multimodal_args(parser: argparse.ArgumentParser) -> None:
    parser.add_argument("--multimodal-translation-config-yaml", type=str, default=None,
                        help="YAML with the fusion keys (SA_image_dropout, use_selective_gate, image_feat_dim, ...)")
    parser.add_argument("--freezing-updates", type=int, default=None)


def _cfg_attr(cfg, attr, default=None):
    """The reference's ``get_attr_from_config`` (tasks/speech_to_speech.py:95-100) on our YAML dict."""
    if cfg is None:
        return default
    v = cfg.get(attr, None)
    return default if v is None else v


try:  # fairseq is not installed in the build image: exercised with a stub in tests/test_host_fairseq_stub.py
    from fairseq.tasks import register_task
    from fairseq.tasks.speech_to_speech import SpeechToSpeechTask

    @register_task(TASK_NAME)
    class MultiModalSpeechToSpeechTask(SpeechToSpeechTask):
        @classmethod
        def add_args(cls, parser):
            super().add_args(parser)
            add_multimodal_args(parser)

        def load_dataset(self, split, epoch=1, combine=False, **kwargs):
            """Same call as the reference's task (tasks/speech_to_speech.py:102-128): the dataset -- manifests, image
            feature tensors, target units, the collater that puts ``imgs_list`` / ``img_masks_list`` into
            ``net_input`` -- is host-side I/O outside this path and stays with the reference's own data package, which
            must be importable (``PYTHONPATH=<reference root>``)."""
            try:
                from mm_s2ut.data.speech_to_speech_dataset import MultiModalSpeechToSpeechDatasetCreator
            except ImportError as e:
                raise ImportError(
                    "task multimodal_speech_to_speech builds its dataset with the reference's data package "
                    "(mm_s2ut.data.speech_to_speech_dataset): put the reference root on PYTHONPATH next to "
                    "--user-dir multimodal-s2ut_b200 (INTEGRATION.md section 1)") from e
            from ..config import load_mm_config

            mm = load_mm_config(getattr(self.args, "multimodal_translation_config_yaml", None))
            self.datasets[split] = MultiModalSpeechToSpeechDatasetCreator.from_tsv(
                root=self.args.data, data_cfg=self.data_cfg, splits=split, is_train_split=split.startswith("train"),
                epoch=epoch, seed=self.args.seed, target_is_code=self.args.target_is_code,
                tgt_dict=self.target_dictionary, n_frames_per_step=self.args.n_frames_per_step,
                multitask=self.multitask_tasks, noise_wav=[], noise_prob=0.0, noise_snr=0.0, noise_num=0,
                image_feat_path=_cfg_attr(mm, "image_feat_path"), flickr30k_root=_cfg_attr(mm, "flickr30k_root"),
                load_visual_extractor_type=_cfg_attr(mm, "load_visual_extractor_type"),
                load_visual_extractor=_cfg_attr(mm, "load_visual_extractor"),
                image_input_size=_cfg_attr(mm, "image_input_size"), image_mean=_cfg_attr(mm, "image_mean"),
                image_std=_cfg_attr(mm, "image_std"))
except ImportError:
    MultiModalSpeechToSpeechTask = None
