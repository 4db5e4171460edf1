"""Architecture arguments and the multimodal-translation YAML.

Mirrors the two configuration layers of the reference:

* fairseq argparse args filled in by ``s2ut_architecture_base`` (called from
  ``mm_s2ut_architecture_base`` at mm_s2ut/models/mm_s2s_transformer.py:703-707);
* the YAML read inside the encoder constructor through
  ``args.multimodal_translation_config_yaml`` (mm_s2s_transformer.py:103-107), whose keys are
  listed in mm_s2ut/config/multimodal_s2ut_transformer.yaml:1-41.  The reference loads it with
  OmegaConf<2.1, where a missing key reads as ``None``; ``MMConfig`` reproduces that.
"""
from __future__ import annotations

import argparse
from pathlib import Path
from typing import Any, Dict, Optional, Union

import yaml


def s2ut_architecture_base(args: argparse.Namespace) -> argparse.Namespace:
    """Defaults of fairseq ``s2ut_architecture_base`` + ``s2t``-style encoder defaults."""
    d = args.__dict__
    d.setdefault("input_feat_per_channel", 80)
    d.setdefault("input_channels", 1)
    d.setdefault("conv_kernel_sizes", "5,5")
    d.setdefault("conv_channels", 1024)
    d.setdefault("encoder_embed_dim", 512)
    d.setdefault("encoder_ffn_embed_dim", 2048)
    d.setdefault("encoder_layers", 12)
    d.setdefault("encoder_attention_heads", 8)
    d.setdefault("encoder_normalize_before", True)
    d.setdefault("no_scale_embedding", False)
    d.setdefault("max_source_positions", 6000)
    d.setdefault("dropout", 0.1)
    d.setdefault("attention_dropout", d["dropout"])
    d.setdefault("activation_dropout", d["dropout"])
    d.setdefault("activation_fn", "relu")
    d.setdefault("encoder_freezing_updates", 0)
    d.setdefault("decoder_embed_dim", d["encoder_embed_dim"])
    d.setdefault("decoder_ffn_embed_dim", d["encoder_ffn_embed_dim"])
    d.setdefault("decoder_layers", 6)
    d.setdefault("decoder_attention_heads", 8)
    d.setdefault("decoder_normalize_before", True)
    d.setdefault("share_decoder_input_output_embed", True)
    d.setdefault("target_code_size", 1000)
    d.setdefault("speaker_embed_dim", 256)
    d.setdefault("target_speaker_embed", False)
    d.setdefault("multimodal_translation_config_yaml", None)
    return args


PRESETS: Dict[str, Dict[str, Any]] = {
    # BASELINE.json configs[0]: 6 enc / 6 dec layers, d=256 (fairseq s2ut_transformer_fisher width)
    "small": dict(encoder_embed_dim=256, encoder_ffn_embed_dim=2048, encoder_layers=6,
                  encoder_attention_heads=4, decoder_layers=6, decoder_attention_heads=4),
    # configs[1]: s2ut_architecture_base as shipped
    "base": dict(),
    # configs[4]: 16 enc layers, d=1024
    "large": dict(encoder_embed_dim=1024, encoder_ffn_embed_dim=4096, encoder_layers=16,
                  encoder_attention_heads=16, decoder_attention_heads=16),
}


def make_args(preset: str = "base", **overrides) -> argparse.Namespace:
    ns = argparse.Namespace(**PRESETS[preset])
    ns.__dict__.update(overrides)
    return s2ut_architecture_base(ns)


class MMConfig(dict):
    """Dict with attribute access; absent keys read as ``None`` (OmegaConf<2.1 behaviour)."""

    def __getattr__(self, k):
        return self.get(k, None)

    def __setattr__(self, k, v):
        self[k] = v


DEFAULT_YAML = Path(__file__).parent / "config" / "multimodal_s2ut_transformer.yaml"


def load_mm_config(src: Union[str, Path, Dict[str, Any], None]) -> Optional[MMConfig]:
    if src is None:
        return None
    if isinstance(src, dict):
        return MMConfig(src)
    with open(src, "r") as f:
        return MMConfig(yaml.safe_load(f) or {})
