"""``mm_s2ut_transformer``: host-side mirror of the reference encoder/model interface.

Reference: mm_s2ut/models/mm_s2s_transformer.py -- ``MM_S2STransformerEncoder`` (:87-622, plain-S2T
branch :463-464 and fusion-at-top branch :471-472, :496-530, :557-560, ``fuse_img_feat`` :594-622),
``MM_S2UTTransformerModel`` (:625-700) and the arch function (:703-707).  Same constructor argument
(``args`` namespace), same ``forward`` signature, same encoder-out dict, same ``state_dict`` keys and
YAML keys.  The arithmetic is NOT PyTorch: ``forward`` hands device pointers to the sm_100a kernels
through the C-ABI library (``engine.EncoderEngine``); there is no CPU or eager fallback.

Two deliberate extensions of the input contract:
* ``src_tokens`` may be the RAW waveform ``[B, N]`` float32 (already x 2**15, i.e. int16 range, as the
  reference's ``get_waveform`` produces it at mm_s2ut/data/audio_utils.py:289-290) with ``src_lengths``
  in samples; fbank + utterance CMVN then run on the device.  A 3-D ``[B, T, 80]`` input keeps the
  reference behaviour (features computed upstream).
* a batch without padding yields an all-False mask instead of the reference's IndexError (:527).
"""
from __future__ import annotations

import logging
import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from ..config import load_mm_config, s2ut_architecture_base
from .modules import (
    Conv1dAdaptorParams,
    Conv1dSubsampler,
    Linear,
    MultimodalAttentionParams,
    SelectiveAttentionParams,
    SinusoidalPositionalEmbedding,
    TransformerEncoderLayerParams,
)

logger = logging.getLogger(__name__)

try:  # registered with fairseq when it is importable (--user-dir contract), plain nn.Module otherwise
    from fairseq.models import FairseqEncoder, register_model, register_model_architecture  # type: ignore

    _HAVE_FAIRSEQ = True
except Exception:  # fairseq is not installed in the build image
    FairseqEncoder = None
    _HAVE_FAIRSEQ = False


class S2TTransformerEncoderParams(nn.Module):
    """Parameters of fairseq ``S2TTransformerEncoder`` (the base class the reference subclasses, :87)."""

    def __init__(self, args):
        # nn.Module.__init__ explicitly, not super(): when fairseq is present the registered encoder class also
        # inherits FairseqEncoder (models/fairseq_glue.py), whose __init__ takes a dictionary -- a cooperative
        # super().__init__() here would land there without one
        nn.Module.__init__(self)
        s2ut_architecture_base(args)
        self.args = args
        d = args.encoder_embed_dim
        self.embed_dim = d
        self.num_heads = args.encoder_attention_heads
        self.ffn_dim = args.encoder_ffn_embed_dim
        self.num_layers = args.encoder_layers
        self.embed_scale = 1.0 if args.no_scale_embedding else math.sqrt(d)
        self.padding_idx = 1
        self.dropout_p = args.dropout
        # fairseq: activation_dropout, falling back to relu_dropout when it is 0; attention_dropout inside the MHA
        self.activation_dropout_p = float(getattr(args, "activation_dropout", 0.0) or getattr(args, "relu_dropout", 0.0) or 0.0)
        self.attention_dropout_p = float(getattr(args, "attention_dropout", 0.0) or 0.0)
        self.subsample = Conv1dSubsampler(
            args.input_feat_per_channel * args.input_channels, args.conv_channels, d,
            [int(k) for k in args.conv_kernel_sizes.split(",")])
        self.embed_positions = SinusoidalPositionalEmbedding(d, self.padding_idx, args.max_source_positions + 2)
        self.transformer_layers = nn.ModuleList(
            TransformerEncoderLayerParams(d, self.ffn_dim, self.num_heads) for _ in range(self.num_layers))
        self.layer_norm = nn.LayerNorm(d, eps=1e-5) if args.encoder_normalize_before else None
        if not args.encoder_normalize_before:
            raise NotImplementedError("post-LN S2T encoders are outside the mm_s2ut_transformer arch (pre-LN)")
        if args.activation_fn != "relu":
            raise NotImplementedError("s2ut_architecture_base uses ReLU")


class MM_S2STransformerEncoder(S2TTransformerEncoderParams):
    """S2T transformer encoder + image fusion at the top (SelectiveAttention / MultimodalAttention,
    sigmoid selective gate, modality dropout)."""

    def __init__(self, args, build_unused_projections: bool = True):
        super().__init__(args)
        self.spk_emb_proj = None
        if getattr(args, "target_speaker_embed", False):
            raise NotImplementedError("target_speaker_embed is off on the mm_s2ut_transformer hot path")
        self.multimodal_translation_flag = False
        self.is_fusion_top = False
        self.load_visual_extractor_type = None
        self.only_img = False
        self.multimodal_attention_type = None
        cfg = load_mm_config(getattr(args, "multimodal_translation_config_yaml", None))
        self.mm_config = cfg
        if cfg is not None:
            self.multimodal_translation_flag = True
            self.load_visual_extractor_type = cfg.load_visual_extractor_type
            self.only_img = cfg.only_img
        logger.info(f"only_img = {self.only_img}")
        logger.info(f"load_visual_extractor_type = {self.load_visual_extractor_type}")
        if self.load_visual_extractor_type not in (None, ""):
            raise NotImplementedError("in-model visual extractors are out of scope: image features are precomputed")
        if self.only_img:
            raise NotImplementedError("only_img is an ablation outside the hot path")
        if self.multimodal_translation_flag:
            d = self.embed_dim
            self.multimodal_attention_type = cfg.multimodal_attention_type
            self.use_selective_gate = cfg.use_selective_gate
            logger.info(f"multimodal_attention_type = {self.multimodal_attention_type}")
            logger.info(f"use_selective_gate = {self.use_selective_gate}")
            dims = list(cfg.image_feat_dim)
            if self.multimodal_attention_type is None:
                pass
            elif self.multimodal_attention_type == "selective_attention":
                self.selective_attns = nn.ModuleList(
                    SelectiveAttentionParams(qdim=d, kdim=i, vdim=i, attn_dim=d, intermediate_dim=d, output_dim=d,
                                             num_heads=1, attn_drop=cfg.SA_attention_dropout) for i in dims)
            elif self.multimodal_attention_type == "multimodal_attention":
                self.is_merge_text_img = bool(cfg.is_merge_text_img)
                if self.is_merge_text_img:
                    raise NotImplementedError("is_merge_text_img=True is not on the shipped path")
                self.multimodal_attns = nn.ModuleList(
                    MultimodalAttentionParams(embed_dim=d, kdim=i, vdim=i, num_heads=1,
                                              dropout=cfg.SA_attention_dropout, add_bias_kv=True) for i in dims)
            else:
                raise NotImplementedError(self.multimodal_attention_type)
            self.gate_denses = nn.ModuleList(Linear(2 * d, d) for _ in dims)
            self.SA_image_dropout = float(cfg.SA_image_dropout or 0.0)
            self.SA_text_dropout = float(cfg.SA_text_dropout or 0.0)
            self.SA_attention_dropout = float(cfg.SA_attention_dropout or 0.0)
            self.image_pre_norm_module = nn.Identity()
            if cfg.image_pre_norm:
                # the reference passes the whole list as normalized_shape (:190): one shared LayerNorm
                self.image_pre_norm_module = nn.LayerNorm(dims, 1e-5, True)
            self.is_fusion_top = bool(cfg.is_fusion_top)
            self.modality_dropout = cfg.modality_dropout if cfg.modality_dropout is not None else -1.0
            self.audio_dropout = cfg.audio_dropout if cfg.audio_dropout is not None else -1.0
        self.mhubert_flag = False
        self.wav2vec2_flag = False
        if build_unused_projections:  # always constructed by the reference (:212-224): checkpoint compatibility
            self.proj_768_to_512 = Linear(768, 512)
            self.proj_1024_to_512 = Linear(1024, 512)
            self.proj_1024_to_768 = Linear(1024, 768)
            self.wav2vec2_adaptor = Conv1dAdaptorParams(1024, 768, 3, 3, 2, True)
        self.num_updates = None
        self.freezing_updates = getattr(args, "freezing_updates", None)
        self.modality_rng = np.random  # the reference draws from the global numpy RNG (:497)
        # train-time SpecAugment of the reference's data config, moved to the device with fbank + CMVN
        # (data.specaugment.SpecAugmentTransform or None); set e.g. enc.specaugment = SpecAugmentTransform.from_policy("lb")
        self.specaugment = None
        self._engine = None
        logger.info(f"multimodal_translation_flag = {self.multimodal_translation_flag}")
        logger.info(f"is_fusion_top = {self.is_fusion_top}")

    # -- engine plumbing ---------------------------------------------------------------------------
    def engine(self):
        from ..engine import EncoderEngine

        if self._engine is None:
            self._engine = EncoderEngine(self)
        return self._engine

    def train_engine(self):
        """Engine of the training-step variant (activations kept, backward kernels, flat fp32 parameters)."""
        from ..training import TrainEngine

        if not isinstance(self._engine, TrainEngine):
            self._engine = TrainEngine(self, getattr(self, "op_dtype", None))
        return self._engine

    def _apply(self, fn, *a, **k):  # .to()/.cuda()/.half(): packed device weights are stale
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def set_num_updates(self, num_updates):
        self.num_updates = num_updates

    def _frozen(self) -> bool:
        """fairseq ``S2TTransformerEncoder.forward``: the encoder runs under ``torch.no_grad()`` while
        ``num_updates < encoder_freezing_updates``."""
        n = int(getattr(self.args, "encoder_freezing_updates", 0) or 0)
        return n > 0 and self.num_updates is not None and self.num_updates < n

    def max_positions(self):
        return self.args.max_source_positions

    # -- forward -----------------------------------------------------------------------------------
    def forward(self, src_tokens, src_lengths, src_audio_path=None, img_path=None, img_tensor=None,
                imgs_list=[], img_masks_list=[], tgt_speaker=None, return_all_hiddens=False, **kwargs):
        eng = self.engine()
        fuse = self.multimodal_translation_flag and self.is_fusion_top and bool(imgs_list)
        drop_audio = drop_image = False
        if fuse and self.training:
            p_mod, p_aud = self.modality_rng.random(), self.modality_rng.random()  # per batch (:497)
            if p_mod < self.modality_dropout:
                if p_aud < self.audio_dropout:
                    drop_audio = True        # reference: NameError at :500; intent = zero the speech states
                else:
                    drop_image = True        # (:504-505) every image tensor zeroed
        frozen = self.training and self._frozen()
        if self.training and (torch.is_grad_enabled() or frozen):
            # training step: autograd reaches the path through one Function on the fused states
            from ..training import EncoderOutGrad

            eng = self.train_engine()
            if getattr(self, "_external_optimizer", True):
                eng.refresh_operands()      # a torch / fairseq optimizer may have stepped the fp32 parameters
            if eng.grads_attached():        # on this path autograd owns .grad (accumulation, hooks, DDP reducer)
                for p in eng.params:
                    p.grad = None
            specaug = None
            if self.specaugment is not None:
                from ..engine import num_frames

                raw = src_tokens.dim() == 2
                frames = [num_frames(int(n)) if raw else int(n) for n in src_lengths.tolist()]   # host sync: training only
                sa = self.specaugment
                specaug = (torch.from_numpy(sa.draw_batch(frames, 80, self.modality_rng)), sa.freq_mask_n,
                           sa.time_mask_n, sa.mask_value)
            out = eng.forward_train(src_tokens, src_lengths, imgs_list if fuse else [], img_masks_list if fuse else [],
                                    drop_audio=drop_audio, drop_image=drop_image, specaug=specaug,
                                    return_all_hiddens=return_all_hiddens)
            if not frozen:      # inside fairseq's encoder_freezing_updates window the states carry no gradient
                out["encoder_out"] = [EncoderOutGrad.apply(out["encoder_out"][0], eng, eng.generation, *eng.params)]
            return out
        from ..training import TrainEngine

        if isinstance(eng, TrainEngine) and getattr(self, "_external_optimizer", True):
            # validation / generation after training with a torch or fairseq optimizer: that optimizer stepped the fp32
            # parameters, the 16-bit operand copies (and the re-laid-out conv weights) are one step old
            eng.refresh_operands()
        return eng.forward(src_tokens, src_lengths, imgs_list if fuse else [], img_masks_list if fuse else [],
                           return_all_hiddens=return_all_hiddens, drop_audio=drop_audio, drop_image=drop_image,
                           training=self.training)

    def reorder_encoder_out(self, encoder_out: Dict[str, List[torch.Tensor]], new_order):
        """fairseq ``S2TTransformerEncoder.reorder_encoder_out`` (beam search)."""
        return {
            "encoder_out": [x.index_select(1, new_order) for x in encoder_out["encoder_out"]],
            "encoder_padding_mask": [x.index_select(0, new_order) for x in encoder_out["encoder_padding_mask"]],
            "encoder_embedding": [x.index_select(0, new_order) for x in encoder_out["encoder_embedding"]],
            "encoder_states": [x.index_select(1, new_order) for x in encoder_out["encoder_states"]],
            "src_tokens": [],
            "src_lengths": [],
        }


def mm_s2ut_architecture_base(args):
    """Arch ``mm_s2ut_transformer`` = fairseq ``s2ut_architecture_base`` (reference :703-707)."""
    return s2ut_architecture_base(args)


if _HAVE_FAIRSEQ:  # pragma: no cover - exercised only where fairseq is installed
    from .fairseq_glue import register as _register

    _register(MM_S2STransformerEncoder, mm_s2ut_architecture_base)
