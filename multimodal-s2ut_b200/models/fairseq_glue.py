"""fairseq registration (only imported when fairseq is installed; the build image has no fairseq, so this module
is exercised on the user's side).  Mirrors mm_s2ut/models/mm_s2s_transformer.py:625-707 of the reference:
``@register_model("mm_s2ut_transformer")`` on an ``S2UTTransformerModel`` subclass whose ``build_encoder`` returns
our encoder, plus ``@register_model_architecture("mm_s2ut_transformer", "mm_s2ut_transformer")``."""
from __future__ import annotations

import logging
from pathlib import Path

logger = logging.getLogger(__name__)


def register(encoder_cls, arch_fn):
    from fairseq import checkpoint_utils
    from fairseq.models import FairseqEncoder, register_model, register_model_architecture
    from fairseq.models.speech_to_speech.s2s_transformer import S2UTTransformerModel, s2ut_architecture_base

    class _Encoder(encoder_cls, FairseqEncoder):  # FairseqEncoder supplies the generator-facing helpers
        def __init__(self, args):
            # encoder_cls initialises nn.Module itself (explicitly, see S2TTransformerEncoderParams.__init__), so the
            # MRO never routes an argument-less __init__ into FairseqEncoder.__init__(dictionary); what that
            # constructor would have done for a speech encoder is one attribute:
            encoder_cls.__init__(self, args)
            self.dictionary = None

    @register_model("mm_s2ut_transformer")
    class MM_S2UTTransformerModel(S2UTTransformerModel):
        @classmethod
        def build_encoder(cls, args):
            encoder = _Encoder(args)
            path = getattr(args, "load_pretrained_encoder_from", None)
            if path is not None:
                if not Path(path).exists():
                    logger.warning(f"skipped pretraining because {path} does not exist")
                else:
                    encoder = checkpoint_utils.load_pretrained_component_from_model(component=encoder, checkpoint=path)
                    logger.info(f"loaded pretrained encoder from: {path}")
            return encoder

        def forward_encoder(self, src_tokens, src_lengths, src_audio_path, img_path, img_tensor, imgs_list,
                            img_masks_list, speaker=None, **kwargs):
            return self.encoder(src_tokens, src_lengths=src_lengths, src_audio_path=src_audio_path,
                                img_path=img_path, img_tensor=img_tensor, imgs_list=imgs_list,
                                img_masks_list=img_masks_list, tgt_speaker=speaker, **kwargs)

        def forward(self, src_tokens, src_lengths, prev_output_tokens, src_audio_path, img_path, img_tensor,
                    imgs_list=[], img_masks_list=[], tgt_speaker=None, return_all_hiddens=False, **kwargs):
            encoder_out = self.forward_encoder(src_tokens, src_lengths=src_lengths, src_audio_path=src_audio_path,
                                               img_path=img_path, img_tensor=img_tensor, imgs_list=imgs_list,
                                               img_masks_list=img_masks_list, speaker=tgt_speaker,
                                               return_all_hiddens=return_all_hiddens, **kwargs)
            decoder_out = self.decoder(prev_output_tokens, encoder_out=encoder_out)
            if return_all_hiddens:
                decoder_out[-1]["encoder_states"] = encoder_out["encoder_states"]
                decoder_out[-1]["encoder_padding_mask"] = encoder_out["encoder_padding_mask"]
            return decoder_out

    @register_model_architecture(model_name="mm_s2ut_transformer", arch_name="mm_s2ut_transformer")
    def mm_s2ut_architecture_base(args):
        s2ut_architecture_base(args)

    return MM_S2UTTransformerModel
