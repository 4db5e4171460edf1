"""The fairseq side of the drop-in boundary, executed against a STUB fairseq (fairseq itself is not installed in the
build image): registry names of the reference (``@register_model("mm_s2ut_transformer")`` + the architecture of the
same name, mm_s2ut/models/mm_s2s_transformer.py:625, :703-707; ``@register_task("multimodal_speech_to_speech")`` with
``--multimodal-translation-config-yaml``, mm_s2ut/tasks/speech_to_speech.py:45-56), the encoder class fairseq would
construct (it must also be a ``FairseqEncoder``), and the task's dataset call."""
import argparse
import importlib.util
import sys
import types
from pathlib import Path

import pytest
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture
def fairseq_stub(monkeypatch):
    reg = {"models": {}, "archs": {}, "tasks": {}}

    class FairseqEncoder(nn.Module):          # same constructor contract as fairseq's
        def __init__(self, dictionary):
            super().__init__()
            self.dictionary = dictionary

        def max_positions(self):
            return 1e6

    class S2UTTransformerModel(nn.Module):
        def __init__(self, encoder, decoder):
            super().__init__()
            self.encoder, self.decoder = encoder, decoder

    class SpeechToSpeechTask:
        def __init__(self, args):
            self.args, self.datasets = args, {}
            self.data_cfg, self.target_dictionary, self.multitask_tasks = "cfg", "dict", {}

        @classmethod
        def add_args(cls, parser):
            parser.add_argument("data")
            parser.add_argument("--target-is-code", action="store_true")

    def register_model(name):
        def deco(cls):
            reg["models"][name] = cls
            return cls
        return deco

    def register_model_architecture(model_name, arch_name):
        def deco(fn):
            reg["archs"][(model_name, arch_name)] = fn
            return fn
        return deco

    def register_task(name):
        def deco(cls):
            reg["tasks"][name] = cls
            return cls
        return deco

    mods = {n: types.ModuleType(n) for n in (
        "fairseq", "fairseq.models", "fairseq.checkpoint_utils", "fairseq.models.speech_to_speech",
        "fairseq.models.speech_to_speech.s2s_transformer", "fairseq.tasks", "fairseq.tasks.speech_to_speech")}
    mods["fairseq"].checkpoint_utils = mods["fairseq.checkpoint_utils"]
    mods["fairseq.models"].FairseqEncoder = FairseqEncoder
    mods["fairseq.models"].register_model = register_model
    mods["fairseq.models"].register_model_architecture = register_model_architecture
    mods["fairseq.models.speech_to_speech.s2s_transformer"].S2UTTransformerModel = S2UTTransformerModel
    mods["fairseq.models.speech_to_speech.s2s_transformer"].s2ut_architecture_base = lambda args: None
    mods["fairseq.tasks"].register_task = register_task
    mods["fairseq.tasks.speech_to_speech"].SpeechToSpeechTask = SpeechToSpeechTask
    for n, m in mods.items():
        monkeypatch.setitem(sys.modules, n, m)
    reg["FairseqEncoder"] = FairseqEncoder
    return reg


def test_model_registration_and_encoder_construction(fairseq_stub):
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200.config import DEFAULT_YAML, make_args
    from mm_s2ut_b200.models import fairseq_glue
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder, mm_s2ut_architecture_base

    Model = fairseq_glue.register(MM_S2STransformerEncoder, mm_s2ut_architecture_base)
    assert fairseq_stub["models"]["mm_s2ut_transformer"] is Model
    assert ("mm_s2ut_transformer", "mm_s2ut_transformer") in fairseq_stub["archs"]
    args = make_args("small", multimodal_translation_config_yaml=str(DEFAULT_YAML))
    torch.manual_seed(0)
    enc = Model.build_encoder(args)                      # what fairseq's build_model calls
    assert isinstance(enc, fairseq_stub["FairseqEncoder"]) and isinstance(enc, MM_S2STransformerEncoder)
    assert enc.dictionary is None
    torch.manual_seed(0)
    plain = MM_S2STransformerEncoder(make_args("small", multimodal_translation_config_yaml=str(DEFAULT_YAML)))
    assert list(enc.state_dict().keys()) == list(plain.state_dict().keys())
    assert all(torch.equal(a, b) for (k, a), b in zip(enc.state_dict().items(), plain.state_dict().values())
               if k != "embed_positions._float_tensor")       # fairseq's uninitialised 1-element buffer
    # the reference's checkpoint names are there
    keys = set(enc.state_dict().keys())
    for k in ("subsample.conv_layers.0.weight", "transformer_layers.0.self_attn.q_proj.weight", "layer_norm.weight",
              "selective_attns.0.q_proj.weight", "gate_denses.0.weight", "image_pre_norm_module.weight",
              "proj_768_to_512.weight", "wav2vec2_adaptor.layers.0.weight"):
        assert k in keys, k
    assert enc.max_positions() == args.max_source_positions   # our method wins over FairseqEncoder's in the MRO


def test_task_registration_flag_and_dataset_call(fairseq_stub, monkeypatch):
    import mm_s2ut_b200  # noqa: F401

    path = ROOT / "multimodal-s2ut_b200" / "tasks" / "speech_to_speech.py"
    spec = importlib.util.spec_from_file_location("mm_s2ut_b200.tasks._under_stub", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    Task = fairseq_stub["tasks"]["multimodal_speech_to_speech"]
    assert Task is mod.MultiModalSpeechToSpeechTask
    parser = argparse.ArgumentParser()
    Task.add_args(parser)
    from mm_s2ut_b200.config import DEFAULT_YAML

    a = parser.parse_args(["/data", "--multimodal-translation-config-yaml", str(DEFAULT_YAML), "--target-is-code"])
    assert a.multimodal_translation_config_yaml == str(DEFAULT_YAML)
    a.seed, a.n_frames_per_step = 1, 1
    task = Task(a)
    with pytest.raises(ImportError, match="reference's data package"):      # no silent net_input without imgs_list
        task.load_dataset("train")
    calls = {}
    ds = types.ModuleType("mm_s2ut.data.speech_to_speech_dataset")

    class Creator:
        @staticmethod
        def from_tsv(**kw):
            calls.update(kw)
            return "dataset"

    ds.MultiModalSpeechToSpeechDatasetCreator = Creator
    for n in ("mm_s2ut", "mm_s2ut.data"):
        monkeypatch.setitem(sys.modules, n, types.ModuleType(n))
    monkeypatch.setitem(sys.modules, "mm_s2ut.data.speech_to_speech_dataset", ds)
    task.load_dataset("train_mm")
    assert task.datasets["train_mm"] == "dataset"
    assert calls["root"] == "/data" and calls["is_train_split"] and calls["target_is_code"]
    assert "image_feat_path" in calls and calls["load_visual_extractor_type"] is None


def test_encoder_freezing_window_routes_to_the_no_grad_forward():
    """fairseq S2TTransformerEncoder: no gradient while num_updates < encoder_freezing_updates."""
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200.config import DEFAULT_YAML, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder

    enc = MM_S2STransformerEncoder(make_args("small", multimodal_translation_config_yaml=str(DEFAULT_YAML),
                                             encoder_freezing_updates=100), build_unused_projections=False)
    assert not enc._frozen()                 # num_updates unknown: fairseq's behaviour is "not frozen"
    enc.set_num_updates(10)
    assert enc._frozen()
    enc.set_num_updates(100)
    assert not enc._frozen()
