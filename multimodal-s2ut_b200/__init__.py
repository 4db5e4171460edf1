"""mm-s2ut-b200: B200-native (sm_100a) fbank -> fused-encoder hot path of VisualTrans.

Drop-in for the ``mm_s2ut_transformer`` arch / ``multimodal_speech_to_speech`` task of
whxhcj/multimodal-S2UT.  The directory doubles as a fairseq ``--user-dir`` (fairseq auto-imports
``models/`` and ``tasks/``); outside fairseq use ``import mm_s2ut_b200`` (alias module at the repo root).
"""
__version__ = "0.1.0"
